"""N>1 host logic on CPU: world_size-2 gloo processes exercise deeptables_b200/dp.py (the exact
functions DeepModel.train_step calls between backward and Adam)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from deeptables_b200 import dp
        assert dp.is_distributed() and dp.world_size() == world
        g = torch.Generator().manual_seed(100 + rank)
        # per-rank "local batch mean" gradients
        dz = torch.ones(4, 1)
        dp.scale_for_mean(dz)
        assert torch.allclose(dz, torch.full((4, 1), 1.0 / world))
        flat = torch.randn(10, generator=g)
        table = torch.zeros(16, 2)
        ids = torch.tensor([[rank, 3 + rank], [5, 7 + rank]], dtype=torch.int32)     # row 5 touched by both
        offs = torch.tensor([0, 6])
        for b in range(2):
            for f in range(2):
                table[offs[f] + ids[b, f]] += 1.0 + rank
        flat_local, table_local = flat.clone(), table.clone()
        union = dp.exchange(flat, table, ids)
        # reference: gather everything and sum
        gathered_flat = [torch.empty(10) for _ in range(world)]
        dist.all_gather(gathered_flat, flat_local)
        gathered_tab = [torch.empty(16, 2) for _ in range(world)]
        dist.all_gather(gathered_tab, table_local)
        assert torch.allclose(flat, sum(gathered_flat))
        assert torch.allclose(table, sum(gathered_tab))
        assert union.shape == (2 * world, 2)
        assert torch.equal(union[2 * rank:2 * rank + 2], ids)
        # union covers every touched row
        touched = (table.abs().sum(dim=1) > 0).nonzero().flatten().tolist()
        rows = sorted({int(offs[f] + union[b, f]) for b in range(union.shape[0]) for f in range(2)})
        assert touched == rows
        # row-wise exchange path (what DeepModel uses on GPU through dtb_grad_rows_pack/unpack),
        # emulated on CPU tensors: owner-of-row semantics, rank-ordered adds
        table2 = table_local.clone()
        flat2 = flat_local.clone()

        def pack():
            packed = torch.zeros(2, 2, 2)
            seen = set()
            for b in range(2):
                for f in range(2):
                    r = int(offs[f] + ids[b, f])
                    if r not in seen:
                        seen.add(r)
                        packed[b, f] = table2[r]
                        table2[r] = 0
            return packed
        order = []

        def unpack(ids_w, packed_w):
            order.append(int(ids_w[0, 0]))
            for b in range(2):
                for f in range(2):
                    table2[int(offs[f] + ids_w[b, f])] += packed_w[b, f]
        union2 = dp.exchange(flat2, table2, ids, pack, unpack)
        assert torch.equal(union2, union) and torch.allclose(table2, table) and torch.allclose(flat2, flat)
        assert order == list(range(world))            # ranks applied in rank order on every replica
        # the split form DeepModel uses to overlap the all-gathers with the CIN weight-gradient kernels:
        # begin (pack + async all-gathers) inside backward, finish (wait + rank-ordered unpack) before Adam
        table2 = table_local.clone()
        order.clear()
        early = dp.TableExchange(ids, pack, unpack)
        assert order == [] and float(table2[int(offs[0] + ids[0, 0])].abs().sum()) == 0.0   # packed out, not yet added back
        union3 = early.finish()
        assert torch.equal(union3, union) and torch.allclose(table2, table)
        assert order == list(range(world))
        w = torch.full((3,), float(rank))
        dp.broadcast_parameters([w, None])
        assert torch.equal(w, torch.zeros(3))
        results[rank] = 'ok'
    finally:
        dist.destroy_process_group()


def test_dp_exchange_world2_gloo():
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, port, results), nprocs=2, join=True)
    assert dict(results) == {0: 'ok', 1: 'ok'}


def test_dp_single_process_is_identity():
    from deeptables_b200 import dp
    assert not dp.is_distributed() and dp.world_size() == 1
    ids = torch.tensor([[1, 2]], dtype=torch.int32)
    g = torch.ones(3)
    assert dp.exchange(g, torch.zeros(2, 2), ids) is ids
    assert torch.equal(dp.scale_for_mean(torch.ones(2)), torch.ones(2))
