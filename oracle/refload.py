"""TEST INFRASTRUCTURE -- not product code.

Imports the *unmodified* reference modules from /root/reference/code so that the
oracle restatement (oracle/sasrec_np.py, oracle/torch_port.py)
can be pinned against the reference itself, and golden vectors can be generated
(oracle/make_golden.py -> tests/golden/*.npz).

/root/reference exists only in the build container; nothing that runs on the GPU
box imports this file.  The reference tree is read-only, hence
sys.dont_write_bytecode.

The reference imports a few packages that are not installed here and that are
not on the hot path: torch_geometric (model/layers.py:9-10,
data/dataload.py:14), colorlog / colorama (utils/logger.py:3,7), tensorboardX
(utils/utils.py:8), lmdb (data/dataset/trainset.py:8), clip (model/load.py:2).
They are replaced by empty stub modules; none of their symbols is executed by
SASRec / GRU4Rec / TransformerEncoder.
"""
import importlib
import os
import sys
import types

REFERENCE_CODE = "/root/reference/code"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_CODE, "REC"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    try:
        import torch_geometric  # noqa: F401
    except Exception:
        class MessagePassing:  # never instantiated on the SASRec path
            def __init__(self, *a, **k):
                pass

        _stub("torch_geometric")
        _stub("torch_geometric.nn", MessagePassing=MessagePassing)
        _stub("torch_geometric.utils", add_self_loops=None, degree=None)
    for name in ("colorlog", "colorama", "tensorboardX", "lmdb", "clip"):
        try:
            importlib.import_module(name)
        except Exception:
            if name == "colorlog":
                _stub(name, ColoredFormatter=object)
            elif name == "colorama":
                _stub(name, init=lambda *a, **k: None)
            elif name == "tensorboardX":
                _stub(name, SummaryWriter=object)
            else:
                _stub(name)


def load_reference():
    """Returns the reference `REC` package (IDNet SASRec / GRU4Rec importable)."""
    if not reference_available():
        raise RuntimeError("/root/reference is not present (GPU box?) -- goldens in tests/golden are the pinned copy")
    sys.dont_write_bytecode = True
    install_stubs()
    if REFERENCE_CODE not in sys.path:
        sys.path.insert(0, REFERENCE_CODE)
    import REC  # noqa: F401
    return REC


class RefDataload:
    """Minimal stand-in for REC.data.dataload.Data: the model only reads item_num (sasrec.py:29)."""

    def __init__(self, item_num, user_num=0):
        self.item_num = item_num
        self.user_num = user_num


def ref_sasrec(cfg: dict, item_num: int):
    """Build the reference SASRec (code/REC/model/IDNet/sasrec.py:9) on CPU from a plain dict config."""
    load_reference()
    from REC.model.IDNet.sasrec import SASRec

    class Cfg(dict):
        def __getitem__(self, k):
            return self.get(k, None)

    return SASRec(Cfg(cfg), RefDataload(item_num))


def ref_gru4rec(cfg: dict, item_num: int):
    load_reference()
    from REC.model.IDNet.gru4rec import GRU4Rec

    class Cfg(dict):
        def __getitem__(self, k):
            return self.get(k, None)

    return GRU4Rec(Cfg(cfg), RefDataload(item_num))
