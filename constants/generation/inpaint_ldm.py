"""Model keys of the inpainting checkpoints (reference: constants/generation/inpaint_ldm.py:1-19)."""
HF_MODEL_KEYS = {
    "realisticvision": "Uminosachi/realisticVisionV51_v51VAE-inpainting",
    "sd-1.5": "runwayml/stable-diffusion-inpainting",
}
