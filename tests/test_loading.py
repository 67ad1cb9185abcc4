"""CPU: checkpoints in the diffusers on-disk layout load into the B200 modules (rtti_b200/loading.py) — the drop-in claim
for real weights (the reference downloads them from the hub: models/region_diffusion.py:24-33, region_diffusion_sdxl.py:105-120)."""
import json
import os

import pytest
import torch

from oracle import unet_oracle as uo
from rtti_b200 import loading
from rtti_b200.vae import AutoencoderKLDecoder, VAEConfig

safetensors = pytest.importorskip("safetensors.torch")


def test_unet_directory_roundtrip(tmp_path):
    cfg = uo.tiny_xl_config()
    sd = {k: v.contiguous() for k, v in uo.make_state_dict(cfg, 3).items()}     # reference (= diffusers) parameter names
    folder = tmp_path / "unet"
    folder.mkdir()
    (folder / "config.json").write_text(json.dumps(cfg.__dict__))
    safetensors.save_file(sd, str(folder / "diffusion_pytorch_model.safetensors"))
    unet = loading.load_unet(str(folder), "cpu")
    got = dict(unet.named_parameters())
    assert len(got) > 50
    # tensors that finalize() does not fold into others survive the trip exactly up to its fp16 cast
    for name in ("conv_in.weight", "time_embedding.linear_1.weight", "mid_block.resnets.0.norm1.weight",
                 "down_blocks.1.attentions.0.transformer_blocks.0.attn2.to_out.0.weight"):
        assert torch.equal(got[name].float(), sd[name].half().float()), name
    del sd["conv_out.bias"]
    safetensors.save_file(sd, str(folder / "diffusion_pytorch_model.safetensors"))
    with pytest.raises(RuntimeError, match="missing"):
        loading.load_unet(str(folder), "cpu")


def test_vae_directory_with_pre_0_18_attention_names(tmp_path):
    cfg = VAEConfig(block_out_channels=(32, 32, 64, 64))
    ref = AutoencoderKLDecoder(cfg).init_synthetic(seed=1)
    sd = {}
    for k, v in ref.state_dict().items():
        for new, old in (("to_q", "query"), ("to_k", "key"), ("to_v", "value"), ("to_out.0", "proj_attn")):
            k = k.replace(f".attentions.0.{new}.", f".attentions.0.{old}.")
        sd[k] = v.contiguous()
    sd["encoder.conv_in.weight"] = torch.zeros(4, 3, 3, 3)      # encoder tensors are ignored
    folder = tmp_path / "vae"
    folder.mkdir()
    (folder / "config.json").write_text(json.dumps({"block_out_channels": [32, 32, 64, 64], "scaling_factor": 0.13025}))
    safetensors.save_file(sd, str(folder / "diffusion_pytorch_model.safetensors"))
    vae = loading.load_vae(str(folder), "cpu", 0.18215)
    assert vae.config.scaling_factor == 0.13025
    for (k, a), (_, b) in zip(sorted(vae.state_dict().items()), sorted(ref.state_dict().items())):
        assert torch.equal(a.float(), b.float()), k


def test_missing_directory_message():
    with pytest.raises(FileNotFoundError, match="no network access"):
        loading.load_sdxl_components("/nonexistent/sdxl", "cpu")
