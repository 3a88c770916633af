"""CPU end-to-end check of the long-sequence attention path INSIDE the ViT item encoder: our CLIP ViT module (model/vit.py) on
the 197-token golden (HF CLIPVisionModel + the reference's MeanItemEncoder, tests/golden/vit_long197.npz), with
  * ops.attention -> ops.LongAttnFn (the real autograd wrapper) calling the REAL device code of csrc/attn_long.cuh /
    attn_long_tc.cuh compiled for the host on the emulation layer (tests/emu), through the same argument lists as the C ABI;
  * the LayerNorm / activation ops, which have run on the GPU and are not under test here, replaced by their torch equivalents.
Pins the Python plumbing (fused qkv layout, head offsets, saved ctx / lse, gradient routing through two trainable layers
and a frozen prefix) together with the kernels' logic at the sequence length of ViT-B/16.  Test infrastructure only."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_P, _I, _LL = C.c_void_p, C.c_int, C.c_longlong


class _EmuLib:
    """stands in for libpixelrec_b200.so for the two long-attention entry points (host pointers, emulated kernels)"""

    def __init__(self, tensor_core):
        out = os.path.join(tempfile.mkdtemp(prefix="pr_emu_"), "libemu.so")
        src = os.path.join(ROOT, "tests", "emu", "emu_kernels.cpp")
        r = subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", out, src], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        self.lib = C.CDLL(out)
        self.lib.emu_attn_long_fwd.argtypes = [_P, _P, _P, _LL, _P, _I, _I, _I, _I, _I, _P, _P, _I]
        self.lib.emu_attn_long_tc_fwd.argtypes = [_P, _P, _P, _LL, _P, _I, _I, _I, _I, _I, _P, _P, _I]
        self.lib.emu_attn_long_bwd.argtypes = [_P, _P, _P, _LL, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _LL, _P, _I]
        self.lib.emu_attn_long_tc_bwd.argtypes = self.lib.emu_attn_long_bwd.argtypes
        self.tc = tensor_core

    def pr_attn_long_fwd_f32(self, q, k, v, ld, key_ids, B, L, h, dh, causal, ctx, lse, stream):
        fn = self.lib.emu_attn_long_tc_fwd if self.tc else self.lib.emu_attn_long_fwd
        fn(q, k, v, ld, key_ids, B, L, h, dh, causal, ctx, lse, 2)
        return 0

    def pr_attn_long_bwd_f32(self, q, k, v, ld, key_ids, ctx, lse, dctx, B, L, h, dh, causal, dq, dk, dv, ld_grad, delta, stream):
        fn = self.lib.emu_attn_long_tc_bwd if self.tc else self.lib.emu_attn_long_bwd
        fn(q, k, v, ld, key_ids, ctx, lse, dctx, B, L, h, dh, causal, dq, dk, dv, ld_grad, delta, 2)
        return 0


def _rel(a, b, floor=1e-7):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), floor))


FULL = os.environ.get("PR_EMU_SLOW") == "1"     # forward + backward on all 3 golden images takes ~100 s of emulation; the default
                                                # run checks the forward on the first image (images are independent) in ~10 s


@pytest.mark.parametrize("tensor_core", [False])      # dh = 16 in this fixture: the mma.sync kernels need dh in {32, 64, 128}
def test_vit_197_tokens_through_the_emulated_long_attention(monkeypatch, tensor_core):
    from pixelrec_b200 import ops
    from pixelrec_b200.model import vit as V
    emu = _EmuLib(tensor_core)
    monkeypatch.setattr(ops, "_L", lambda: emu)
    monkeypatch.setattr(ops, "_req", lambda t, dtype, name: t)
    monkeypatch.setattr(ops, "_stream", lambda t: 0)
    monkeypatch.setattr(V, "_ln", lambda x, ln: F.layer_norm(x, (x.shape[-1],), ln.weight, ln.bias, ln.eps))
    monkeypatch.setattr(ops, "activation", lambda x, name: x * torch.sigmoid(1.702 * x) if name == "quick_gelu" else torch.relu(x))
    z = np.load(os.path.join(ROOT, "tests", "golden", "vit_long197.npz"))
    m = V.CLIPVisionModel(V.CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=4,
                                             image_size=112, patch_size=8))
    m.vision_model.post_layernorm = V.Identity()
    enc = V.MeanItemEncoder(m, 64, 48, "relu")
    enc.load_state_dict({k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")})
    for index, (_, p) in enumerate(m.named_parameters()):
        if index < 21:                                    # embeddings, pre-LN and the first layer frozen, as in the golden
            p.requires_grad = False
    if not FULL:
        with torch.no_grad():
            out = enc(torch.from_numpy(z["x"][:1]))
        assert _rel(out.numpy(), z["out"][:1]) < 1e-4
        return
    out = enc(torch.from_numpy(z["x"]))
    assert out.shape == z["out"].shape
    assert _rel(out.detach().numpy(), z["out"]) < 1e-4
    out.backward(torch.from_numpy(z["gout"]))
    checked = 0
    for k, p in enc.named_parameters():
        if "grad/" + k in z.files:
            assert p.grad is not None, k
            if k.endswith("k_proj.bias"):                 # mathematically zero (softmax shift invariance): fp noise only
                assert p.grad.abs().max().item() < 1e-5
            else:
                assert _rel(p.grad.numpy(), z["grad/" + k], 1e-6) < 1e-3, k
            checked += 1
        else:
            assert p.grad is None, k
    assert checked == 34
