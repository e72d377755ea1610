#!/bin/bash
# experiment: PCIe ceiling with N processes copying at once (one per GPU), next to one process alone
N=${1:-8}
echo "== one process alone (GPU 0)"; CUDA_VISIBLE_DEVICES=0 python tools/exp_pcie.py
echo "== $N processes at once (lines of all processes)"
for g in $(seq 0 $((N-1))); do CUDA_VISIBLE_DEVICES=$g python tools/exp_pcie.py > gpurun_out/pcie8_$g.log 2>&1 & done
wait
grep -h "both, 1 chunk" gpurun_out/pcie8_*.log
grep -h "H2D alone" gpurun_out/pcie8_*.log | head -3
nproc; free -g | head -2
