#!/bin/bash
# Multi-GPU fan-out of the inpainting stage — same role and flags as the reference's scripts/generation/inpaint.sh
# (:72-196 flags, :207-268 one process per GPU with --parallel_idx i --parallel_num N). Here the fan-out is torchrun:
# RANK / WORLD_SIZE become --parallel_idx / --parallel_num, so every rank takes the reference's contiguous work-list slice.
# Every other flag is forwarded to src/generation/inpaint.py, including --no_skip_done, --segmenter module:factory (the
# plug-in human segmenter) and --adaptive_mask_model_type stub.
gpus=()
pass=()
while [[ $# -gt 0 ]]; do
  case $1 in
    --gpus) shift; while [[ $# -gt 0 && $1 != --* ]]; do gpus+=("$1"); shift; done ;;
    --no_skip_done) pass+=("--no_skip_done"); shift ;;
    *) pass+=("$1"); shift ;;
  esac
done
n=${#gpus[@]}; [ "$n" -eq 0 ] && gpus=(0) && n=1
export CUDA_VISIBLE_DEVICES=$(IFS=,; echo "${gpus[*]}")
if [ "$n" -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 src/generation/inpaint.py "${pass[@]}"
else
  python src/generation/inpaint.py "${pass[@]}"
fi
