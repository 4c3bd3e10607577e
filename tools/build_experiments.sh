#!/bin/bash
# Rebuild the library WITH the experiment builds of cin_tc_dgrad_kernel (ablation switches 1-7, see cin_tc.cu): profiling
# only -- the product build (deeptables_b200/build.py, __graft_entry__.build) leaves them out.  -DDTB_FIRST_VERSIONS adds the
# first AFM backward (DTB_AFM_BWD=1) and the first FGCNN filter-gradient kernel (DTB_FGCNN_DW=0).
set -e
cd "$(dirname "$0")/.."
NVCCFLAGS_EXTRA="-DDTB_CIN_EXPERIMENTS -DDTB_FIRST_VERSIONS" python deeptables_b200/build.py --force
