#!/bin/bash
# ComA learning stage — same flags as the reference's scripts/learn_coma.sh:4-37:
#   --IoU_threshold_min X --inlier_num_threshold_min N --dataset_type D --supercategory SC --category C [--no_skip_done]
# plus an optional  --gpus "0 1 2 ..."  : with more than one GPU the extraction runs under torchrun, one process per GPU.
# The filter / downsample stages are the reference's own CPU scripts (open3d, licensed SMPL-X files); they are run
# only if present in this tree. The extraction stage runs on the B200 kernels.
skip_done=true
gpus=""
while [[ $# -gt 0 ]]; do
  case $1 in
    --IoU_threshold_min) IoU_threshold_min="$2"; shift 2 ;;
    --inlier_num_threshold_min) inlier_num_threshold_min="$2"; shift 2 ;;
    --dataset_type) dataset_type="$2"; shift 2 ;;
    --supercategory) supercategory="$2"; shift 2 ;;
    --category) category="$2"; shift 2 ;;
    --gpus) gpus="$2"; shift 2 ;;
    --no_skip_done) skip_done=false; shift 1 ;;
    *) echo "Unknown option: $1"; exit 1 ;;
  esac
done
sd=""; [ "$skip_done" = true ] && sd="--skip_done"
ngpu=$(echo $gpus | wc -w)
if [ "$ngpu" -gt 1 ]; then
  export CUDA_VISIBLE_DEVICES=$(echo $gpus | tr ' ' ',')
  run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $ngpu --master-addr 127.0.0.1 --master-port 29533"
else
  [ -n "$gpus" ] && export CUDA_VISIBLE_DEVICES=$gpus
  run="python"
fi
if [ -f src/coma/filter.py ]; then
  python src/coma/filter.py --IoU_threshold_min $IoU_threshold_min --inlier_num_threshold_min $inlier_num_threshold_min --supercategories $supercategory --categories $category $sd
fi
if [ -f src/coma/downsample_human.py ]; then python src/coma/downsample_human.py $sd; fi
if [ -f src/coma/downsample_objects.py ]; then
  for n in 2048 1500 180; do
    python src/coma/downsample_objects.py --dataset_type $dataset_type --supercategories $supercategory --categories $category --number_of_points $n $sd
  done
fi
for kind in object human occupancy; do
  $run src/coma/extract_coma.py --supercategories $supercategory --categories $category --hyperparams_key "qual:${category}_${kind}" $sd
done
