"""CPU check of the ctypes call sites in pixelrec_b200/ops.py: every wrapper must hand the C ABI the number and types of
arguments include/pixelrec_b200.h declares.  Without a GPU the library itself rejects the call (no driver / its own argument
checks) and the wrapper raises PixelRecB200Error; a wrong argument list would surface earlier as ctypes.ArgumentError or
TypeError.  Nothing is computed here -- the device-pointer checks are bypassed only so that the call reaches ctypes."""
import ctypes

import pytest
import torch


@pytest.fixture
def cpu_calls(monkeypatch):
    from pixelrec_b200 import ops
    monkeypatch.setattr(ops, "_req", lambda t, dtype, name: t)
    monkeypatch.setattr(ops, "_stream", lambda t: 0)
    return ops


def _f(*s):
    return torch.zeros(*s)


def _i(*s):
    return torch.zeros(*s, dtype=torch.int64)


def _long_attn_bwd(o):
    class Ctx:                                   # what LongAttnFn.forward leaves behind
        saved_tensors = (_f(1, 100, 3 * 64), _f(1, 100, 64), _f(2, 100))
        key_ids = None
        cfg = (1, 100, 2, 32, 0)
    return o.LongAttnFn.backward(Ctx, _f(1, 100, 64))


CASES = {
    "attention_long_bwd": _long_attn_bwd,
    "gather_rows": lambda o: o.gather_rows(_f(10, 8), _i(5)),
    "scatter_plan": lambda o: o.ScatterPlan(_i(6), 10, 0),
    "adamw_dense": lambda o: o.adamw_dense(_f(8), _f(8), _f(8), _f(8), 1e-3, 0.9, 0.999, 1e-8, 0.1, 1),
    "adamw_rows": lambda o: o.adamw_rows(_f(4, 8), _f(4, 8), _f(4, 8), None, None, 1e-3, 0.9, 0.999, 1e-8, 0.1, 1),
    "add_ln": lambda o: o.add_ln(_f(6, 16), _f(6, 16), _f(16), _f(16), 1e-12),
    "add_ln_dropout": lambda o: o.add_ln(_f(6, 16), _f(6, 16), _f(16), _f(16), 1e-12, p_pre=0.1, seed=3, stream_pre=1),
    "bpr_loss": lambda o: o.bpr_loss(_f(2, 5, 16), _f(2, 2, 6, 16), _i(2, 5)),
    "activation": lambda o: o.activation(_f(4, 8), "gelu"),
    "attention": lambda o: o.attention(_f(2, 10, 3 * 64), None, 2, causal=True, tf32=False),
    "attention_tf32": lambda o: o.attention(_f(2, 10, 3 * 64), None, 2, causal=True, tf32=True),
    "attention_long": lambda o: o.attention(_f(1, 100, 3 * 64), _i(1, 100), 2, causal=False),
    "score_topk": lambda o: o.score_topk(_f(4, 64), _f(50, 64), 5, _i(3), _i(3)),
    "score_topk_exact": lambda o: o.score_topk_exact(_f(4, 64), _f(50, 64), 5, _i(3), _i(3)),
    "table_norm_max": lambda o: o.table_norm_max(_f(50, 64)),
    "colsum_rows": lambda o: o.colsum_rows(_f(70, 16)),
    "gemm": lambda o: o.gemm(_f(6, 64), _f(8, 64), bias=_f(8), epi=o.GEMM_ACT, act="gelu", want_pre=True),
    "gemm_wgrad_splitk": lambda o: o.gemm(_f(64, 8), _f(64, 12), a_mn=True, b_mn=True, splits=2),
    "gemm_act_bwd_colsum": lambda o: o.gemm(_f(6, 64), _f(64, 8), b_mn=True, aux=_f(6, 8), epi=o.GEMM_ACT_BWD, act="gelu",
                                            want_colsum=True),
    "score_prepare_f16": lambda o: o.score_prepare_f16(_f(50, 64)),
    "score_topk_f16": lambda o: o.score_topk_f16(_f(4, 64), torch.zeros(50, 64, dtype=torch.float16), 5, _i(3), _i(3)),
    "score_ce": lambda o: o.score_ce(_f(4, 64), _f(50, 64), _i(4)),
    "gather_rows_peers": lambda o: o.gather_rows_peers(_i(2), 2, 10, 8, _i(5)),
    "push_rows_peers": lambda o: o.push_rows_peers(_f(5, 8), _i(5), 2, 0, 4, 0, _i(2), _i(2), torch.zeros(2, dtype=torch.int32)),
    "seq_batch_build": lambda o: o.seq_batch_build(_i(7, 11), _i(3), 50, 1),
    "set_seed_device": lambda o: o.set_seed_device(_i(1)),
    "shared_buffer": lambda o: o.SharedBuffer(1024, "cuda:0"),
    "shared_open": lambda o: o.shared_open(b"\\0" * 64, "cuda:0"),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_wrapper_argument_lists_match_the_abi(cpu_calls, name):
    if torch.cuda.is_available():
        pytest.skip("argument-marshalling check for GPU-less hosts; the GPU parity tests cover the real calls")
    from pixelrec_b200.lib import PixelRecB200Error
    try:
        CASES[name](cpu_calls)
        if name == "set_seed_device":            # -DPR_SEED_DEV builds (the default) only record the pointer: nothing to refuse
            cpu_calls.set_seed_device(None)
            return
    except PixelRecB200Error:
        return                                   # reached the library, which refused (no device): the expected outcome
    except (ctypes.ArgumentError, TypeError) as e:      # pragma: no cover
        pytest.fail(f"{name}: call site does not match the C ABI: {e}")
    pytest.fail(f"{name}: the library accepted a call without a CUDA device")
