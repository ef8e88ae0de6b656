"""Host-side behaviour of the drop-in modules that does not need a GPU: constructor signatures, attribute and
state-dict names, checkpoint layout and error behaviour (reference: model_m.py:178-248, trainer.py:39-48,58-62),
and the refusal to run without CUDA (no CPU fallback)."""
import inspect
import os

import pytest
import torch

from oracle import clip_oracle as O

B32 = "openai/clip-vit-base-patch32"


@pytest.fixture(scope="module")
def tiny_clip():
    return O.build_hf_clip(B32, seed=0, vision_layers=1, text_layers=1)


def _sig(fn):
    return [(p.name, p.default) for p in inspect.signature(fn).parameters.values() if p.name != "self" and p.kind != p.KEYWORD_ONLY]


def test_constructor_signatures_match_reference():
    from vlm_clip_b200.adapter.clip_adapter import SharedMHSAttentionAdapter, TextAdapter, VisionAdapter
    from vlm_clip_b200.adapter.peclip import ContextAdapter, SharedAdapter, TextualAdapter
    from vlm_clip_b200.model_m import CLIPWithAdapters
    from vlm_clip_b200.model_t import CLIPAdapter
    from vlm_clip_b200.trainer import CLIPAdapterTrainer

    E = inspect.Parameter.empty
    assert _sig(TextAdapter.__init__) == [("hidden_size", E), ("adapter_size", E)]  # adapter/clip_adapter.py:10
    assert _sig(VisionAdapter.__init__) == [("hidden_size", E), ("adapter_size", E)]  # :137
    assert _sig(SharedMHSAttentionAdapter.__init__) == [("text_input_size", 512), ("image_input_size", 768),
                                                        ("hidden_size", 512), ("num_heads", 8), ("dropout", 0.1)]  # :70-77
    assert _sig(TextualAdapter.__init__) == [("input_dim", E), ("hidden_dim", E)]  # adapter/peclip.py:7
    assert _sig(ContextAdapter.__init__) == [("input_dim", E), ("num_heads", E)]  # :24
    assert _sig(SharedAdapter.__init__) == [("input_dim", E), ("num_heads", E)]  # :38
    assert _sig(CLIPWithAdapters.__init__) == [
        ("clip_model_name", "openai/clip-vit-base-patch32"), ("text_adapter_size", 256), ("vision_adapter_size", 256),
        ("shared_adapter_layers", 2), ("freeze_clip", True), ("use_text_adapter", True), ("use_vision_adapter", True),
        ("use_shared_adapters", True)]  # model_m.py:15-25
    assert _sig(CLIPWithAdapters.forward) == [("input_ids", None), ("attention_mask", None), ("pixel_values", None),
                                              ("return_loss", True)]  # model_m.py:127-129
    assert _sig(CLIPAdapter.__init__) == [("model_name", E), ("alpha", 0.2), ("beta", 0.2), ("bottleneck_dim", 64)]  # model_t.py:38
    assert _sig(CLIPAdapterTrainer.__init__)[:8] == [
        ("model", E), ("train_dataloader", E), ("val_dataloader", None), ("learning_rate", 5e-5), ("weight_decay", 0.01),
        ("warmup_steps", 0), ("max_grad_norm", 1.0), ("output_dir", "./clip_adapter_checkpoints")]  # trainer.py:16-26


def test_state_dict_keys_and_param_count(tiny_clip):
    from vlm_clip_b200.model_m import CLIPWithAdapters

    m = CLIPWithAdapters(clip=tiny_clip, use_shared_adapters=False)
    keys = ["down_project.weight", "down_project.bias", "up_project.weight", "up_project.bias", "layer_norm.weight",
            "layer_norm.bias"]
    assert list(m.text_adapter.state_dict().keys()) == keys  # fixture test_checkpoints/test_adapter.pt
    assert list(m.vision_adapter.state_dict().keys()) == keys
    n = sum(p.numel() for name, p in m.named_parameters() if "adapter" in name)
    assert n == 659712
    assert all(not p.requires_grad for p in m.clip.parameters())
    assert all(p.requires_grad for name, p in m.named_parameters() if "adapter" in name)
    m._unfreeze_clip_parameters()
    assert all(p.requires_grad for p in m.clip.parameters())
    m._freeze_clip_parameters()
    ms = CLIPWithAdapters(clip=tiny_clip, use_shared_adapters=True, shared_adapter_layers=2)
    sk = set(ms.shared_adapters.state_dict().keys())
    for k in ("0.text_proj.weight", "0.image_proj.bias", "0.cross_attn.in_proj_weight", "0.cross_attn.out_proj.weight",
              "1.norm3.bias", "1.mlp.0.weight", "1.mlp.2.bias"):
        assert k in sk  # adapter/clip_adapter.py:79-97
    assert sum(p.numel() for p in ms.shared_adapters[0].parameters()) == 3_809_792  # 3.81 M / layer


def test_checkpoint_roundtrip_and_errors(tiny_clip, tmp_path):
    from vlm_clip_b200.model_m import CLIPWithAdapters

    m = CLIPWithAdapters(clip=tiny_clip, use_shared_adapters=False)
    path = str(tmp_path / "ck" / "adapter.pt")
    m.save_adapter_weights(path)
    blob = torch.load(path)
    assert set(blob.keys()) == {"text_adapter", "vision_adapter"}  # model_m.py:188-195
    m2 = CLIPWithAdapters(clip=tiny_clip, use_shared_adapters=False)
    m2.load_adapter_weights(path)
    for a, b in zip(m.text_adapter.parameters(), m2.text_adapter.parameters()):
        assert torch.equal(a, b)
    with pytest.raises(FileNotFoundError):
        m2.load_adapter_weights(str(tmp_path / "missing.pt"))  # model_m.py:216-217
    m3 = CLIPWithAdapters(clip=tiny_clip, use_text_adapter=False, use_shared_adapters=False)
    with pytest.raises(ValueError, match="Text adapter weights found"):
        m3.load_adapter_weights(path)  # model_m.py:225-226
    m4 = CLIPWithAdapters(clip=tiny_clip, use_shared_adapters=True, shared_adapter_layers=1)
    with pytest.raises(ValueError, match="Shared adapter"):
        m4.load_adapter_weights(path)  # model_m.py:244-245
    m5 = CLIPWithAdapters(clip=tiny_clip, use_text_adapter=False, use_vision_adapter=False, use_shared_adapters=False)
    with pytest.raises(ValueError, match="No adapters enabled"):
        m5.save_adapter_weights(path)  # model_m.py:197-198


def test_reference_checkpoint_layout_loads(tiny_clip, tmp_path):
    """A checkpoint in the reference's layout (golden copy of test_adapter.pt's text half + default vision) loads."""
    from pathlib import Path

    from vlm_clip_b200.model_m import CLIPWithAdapters

    g = torch.load(Path(__file__).parent / "golden" / "adapters.pt")
    m = CLIPWithAdapters(clip=tiny_clip, use_shared_adapters=False)
    blob = {"text_adapter": g["text_adapter"], "vision_adapter": m.vision_adapter.state_dict()}
    for k, shape in g["checkpoint_keys"]["vision_adapter"].items():
        assert tuple(blob["vision_adapter"][k].shape) == shape
    p = tmp_path / "ref_layout.pt"
    torch.save(blob, p)
    m.load_adapter_weights(str(p))
    assert torch.equal(m.text_adapter.down_project.weight, g["text_adapter"]["down_project.weight"])


def test_trainer_selects_adapter_params_and_schedule(tiny_clip, tmp_path):
    from transformers.optimization import get_linear_schedule_with_warmup

    from vlm_clip_b200.model_m import CLIPWithAdapters
    from vlm_clip_b200.trainer import CLIPAdapterTrainer, linear_schedule_multiplier

    m = CLIPWithAdapters(clip=tiny_clip, use_shared_adapters=False)
    tr = CLIPAdapterTrainer(m, [None] * 4, output_dir=str(tmp_path / "out"))
    assert len(tr.trainable_params) == 12 and os.path.isdir(tmp_path / "out")  # trainer.py:36-43
    m_off = CLIPWithAdapters(clip=tiny_clip, use_text_adapter=False, use_vision_adapter=False, use_shared_adapters=False)
    with pytest.raises(ValueError, match="empty parameter list"):
        CLIPAdapterTrainer(m_off, [None], output_dir=str(tmp_path / "out2"))
    w = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([w], lr=1.0)
    sched = get_linear_schedule_with_warmup(opt, num_warmup_steps=3, num_training_steps=10)
    for step in range(12):
        assert abs(opt.param_groups[0]["lr"] - linear_schedule_multiplier(step, 3, 10)) < 1e-12
        opt.step()
        sched.step()


def test_no_cpu_fallback(tiny_clip):
    from vlm_clip_b200 import _native as N
    from vlm_clip_b200.adapter.clip_adapter import TextAdapter
    from vlm_clip_b200.model_m import CLIPWithAdapters

    m = CLIPWithAdapters(clip=tiny_clip, use_shared_adapters=False)
    pix, ids, mask = O.synthetic_batch(2)
    with pytest.raises(N.NativeError):
        m(input_ids=ids, attention_mask=mask, pixel_values=pix)
    with pytest.raises(N.NativeError):
        TextAdapter(512, 256)(torch.randn(2, 77, 512))


def test_track_tv_modules_construct(tiny_clip):
    from vlm_clip_b200 import model_t, model_v
    from vlm_clip_b200.constants import EMOTIONS, get_emotion_descriptions

    d = get_emotion_descriptions()
    assert list(d.keys()) == EMOTIONS and all(len(v) == 5 for v in d.values())  # constants.py:15-75 structure
    a = model_t.VisualAdapter(512, 64)
    assert list(a.state_dict().keys()) == ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"]  # model_t.py:16-18
    v = model_v.EnhancedCLIPAdapter(clip=tiny_clip, device="cpu", vlm_context_extractor=object())
    assert [n for n, _ in v.named_children()][:4] == ["model", "visual_adapter", "text_adapter", "context_adapter"]
    assert len(v.get_trainable_parameters()) == 12  # model_v.py:355-360
    assert v.visual_adapter.fc1.out_features == 192  # config.py:15


def test_full_finetune_host_logic(tmp_path):
    """Config 5 (freeze_clip=False): mode detection, parameter selection of trainer(trainable=...), no CPU fallback."""
    from vlm_clip_b200 import _native as N
    from vlm_clip_b200.finetune import _LAYER_PARAMS, track_m_unused_parameters
    from vlm_clip_b200.model_m import CLIPWithAdapters
    from vlm_clip_b200.trainer import CLIPAdapterTrainer

    clip = O.build_hf_clip(B32, seed=0, vision_layers=1, text_layers=1)
    m = CLIPWithAdapters(clip=clip, freeze_clip=True, use_shared_adapters=False)
    assert not m._full_finetune()
    m._unfreeze_clip_parameters()  # model_m.py:72-75
    assert m._full_finetune()
    # the reference's trainer filters by name whatever requires grad (trainer.py:39-43): still the 12 adapter tensors
    assert len(CLIPAdapterTrainer(m, [None], output_dir=str(tmp_path / "a")).trainable_params) == 12
    tr = CLIPAdapterTrainer(m, [None], output_dir=str(tmp_path / "b"), trainable="all")
    unused = {id(p) for p in track_m_unused_parameters(clip)}
    assert len(unused) == 2
    want = [p for p in m.parameters() if p.requires_grad and id(p) not in unused]
    assert [id(p) for p in tr.trainable_params] == [id(p) for p in want]
    assert sum(p.numel() for p in tr.trainable_params) == sum(p.numel() for p in m.parameters()) - 2 * 768
    with pytest.raises(ValueError, match="trainable"):
        CLIPAdapterTrainer(m, [None], output_dir=str(tmp_path / "c"), trainable="backbone")
    # every parameter of an encoder layer is covered by the hand-written backward
    names = {k.split("encoder.layers.0.")[1] for k, _ in clip.named_parameters() if "vision_model.encoder.layers.0." in k}
    assert names == set(_LAYER_PARAMS)
    pix, ids, mask = O.synthetic_batch(2)
    with pytest.raises(N.NativeError):  # CPU model: the fine-tune towers refuse, they do not fall back to HF
        m(input_ids=ids, attention_mask=mask, pixel_values=pix)
    m._freeze_clip_parameters()
    assert not m._full_finetune()
