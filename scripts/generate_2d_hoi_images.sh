#!/bin/bash
# 2D HOI image generation — same flags as the reference's scripts/generate_2d_hoi_images.sh:5-41:
#   --gpus g0 g1 ... --dataset_type D --supercategory SC --category C [--no_skip_done]
# The rendering / mask-selection / prompt stages are the reference's own (Blender, OpenAI API); they run only if present.
skip_done=true
gpu_ids=()
extra=()
while [[ $# -gt 0 ]]; do
  case $1 in
    --gpus) shift; while [[ $# -gt 0 && $1 != --* ]]; do gpu_ids+=("$1"); shift; done ;;
    --dataset_type) dataset_type="$2"; shift 2 ;;
    --supercategory) supercategory="$2"; shift 2 ;;
    --category) category="$2"; shift 2 ;;
    --no_skip_done) skip_done=false; shift 1 ;;
    --segmenter|--adaptive_mask_model_type|--model_dir) extra+=("$1" "$2"); shift 2 ;;   # forwarded to the inpainting stage
    *) echo "Unknown option: $1"; exit 1 ;;
  esac
done
sd=""; [ "$skip_done" = true ] && sd="--skip_done"
[ -f src/generation/render_objects.py ] && blenderproc run src/generation/render_objects.py --dataset_types $dataset_type --supercategories $supercategory --categories $category $sd
[ -f src/generation/select_mask.py ] && python src/generation/select_mask.py --supercategories $supercategory --categories $category $sd
[ -f src/generation/generate_prompts.py ] && python src/generation/generate_prompts.py --supercategories $supercategory --categories $category $sd
nsd=""; [ "$skip_done" = false ] && nsd="--no_skip_done"
bash scripts/generation/inpaint.sh --supercategories $supercategory --categories $category --gpus ${gpu_ids[@]} $nsd "${extra[@]}"
