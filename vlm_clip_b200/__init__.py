"""vlm_clip_b200 — B200-native (sm_100a) implementation of the CLIP adapter fine-tuning hot path of
Quillboltcode/VLM-CLIP behind the reference's own Python module API.

    from vlm_clip_b200.model_m import CLIPWithAdapters          # reference: model_m.py
    from vlm_clip_b200.trainer import CLIPAdapterTrainer        # reference: trainer.py
    from vlm_clip_b200.adapter.clip_adapter import TextAdapter  # reference: adapter/clip_adapter.py

All arithmetic runs in libvlmclip_b200.so (include/vlmclip.h); there is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
