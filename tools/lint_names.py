"""Undefined-global check for modules whose function bodies only run on a GPU box (no pyflakes in the image).
usage: python tools/lint_names.py file.py ...   -> prints 'file:function: name' for every global name that is
referenced inside a function/class body but defined neither at module level nor in builtins."""
import builtins
import symtable
import sys


def module_names(table):
    return {s.get_name() for s in table.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}


def walk(table, top, path, out):
    for child in table.get_children():
        for s in child.get_symbols():
            if s.is_referenced() and s.is_global() and not s.is_assigned():
                n = s.get_name()
                if n not in top and not hasattr(builtins, n) and n not in ('__file__', '__name__', '__doc__'):
                    out.append(f'{path}:{child.get_name()}: {n}')
        walk(child, top, path, out)


def check(path):
    with open(path) as f:
        src = f.read()
    table = symtable.symtable(src, path, 'exec')
    out = []
    walk(table, module_names(table), path, out)
    return out


if __name__ == '__main__':
    problems = [p for f in sys.argv[1:] for p in check(f)]
    print('\n'.join(problems) if problems else 'no undefined globals')
    sys.exit(1 if problems else 0)
