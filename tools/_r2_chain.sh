./tools/_bin/umma_nsweep 2>&1 | tail -8
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "chain or ghost" 2>&1 | tail -15
for v in 1 0; do echo "CHAIN_MMA=$v"; DCFA_CHAIN_MMA=$v timeout 300 python tools/prof_ops.py --ops dark2.1.pw1+3,dark3.1.pw1+3,dark4.1.pw1+3,up2.m0.g1.pw+2,up2.m0.g2.pw+2,up1.m0.g1.pw+2,up1.m0.g2.pw+2 --iters 4 2>&1 | tail -7; done
DCFA_DW_PARTS=1 timeout 300 python tools/prof_ops.py --ops dark2.1.pw1+3,dark3.1.pw1+3,up2.m0.g1.pw+2 --iters 4 2>&1 | tail -3
