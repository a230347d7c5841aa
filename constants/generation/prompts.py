"""Per-category diffusion overrides used by the inpainting driver (values restated from the reference's
constants/generation/prompts.py:63-93,100-163: only `strength` deviates from the CLI defaults for the shipped
categories). Unknown categories fall back to the CLI defaults instead of raising KeyError."""

SC2DIFFUSERCONFIG = {
    "Chair": {"Lounge Chair / Cafe Chair / Office Chair": {"strength": 1.0, "controlnet_conditioning_scale": 0.0}},
    "motorcycle,bike": {"motorcycle,bike": {"strength": 0.9, "controlnet_conditioning_scale": 0.0}},
    "umbrella": {"umbrella": dict()},
    "frypan": {"frypan": dict()},
    "BEHAVE": {"backpack": {"strength": 0.98}},
    "INTERCAP": {"suitcase": {"strength": 0.98}},
}
SCV2DIFFUSERCONFIG = {sc: {c: dict() for c in cats} for sc, cats in SC2DIFFUSERCONFIG.items()}  # per-view overrides
ALLOWED_VIEWPOINT_AUGMENTATIONS = [", full body", "original"]
