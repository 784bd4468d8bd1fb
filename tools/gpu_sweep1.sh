#!/bin/bash
mkdir -p gpurun_out
V=cilqr_b200/lib/variants
{
echo "== N=20 w16_b12 (168 regs)"; python tools/occ_sweep.py --lib $V/libcilqr_b200_w16_b12.so --horizon 20 --batch 32768 --pads 0,4000,8000,17000,32000
echo "== N=20 w16 (252 regs)"; python tools/occ_sweep.py --lib $V/libcilqr_b200_w16.so --horizon 20 --batch 32768 --pads 8000,17000
echo "== N=20 w16_b16 (128 regs)"; python tools/occ_sweep.py --lib $V/libcilqr_b200_w16_b16.so --horizon 20 --batch 32768 --pads 0,8000
echo "== N=100 w32 (252 regs)"; python tools/occ_sweep.py --lib $V/libcilqr_b200_w32.so --horizon 100 --batch 16384 --pads 0,20000
echo "== N=100 w16 (252 regs)"; python tools/occ_sweep.py --lib $V/libcilqr_b200_w16.so --horizon 100 --batch 16384 --pads 0
} 2>&1 | tee gpurun_out/sweep1.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
