#!/bin/bash
V=cilqr_b200/lib/variants
for v in pt1 pt2 pt3; do python tools/occ_sweep.py --lib $V/libcilqr_b200_$v.so --horizon 20 --batch 8192 --pads 32000,8000,0; done
