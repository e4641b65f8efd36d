"""Import the reference's own model files UNCHANGED (authoring container only) — TEST INFRASTRUCTURE.

`/root/reference/model/x3d.py:13-20` needs pytorchvideo/fvcore, which are not installable here;
`oracle/pv_shim` restates the nine classes it uses.  `/root/reference` does not exist on the GPU
box, so nothing that runs there may call this; it is used by `oracle/make_golden.py` and by the
`not gpu` tests that pin `oracle/change3d_oracle.py` against the real reference (skipped when the
reference tree is absent).
"""
import argparse
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "pv_shim")


def _resolve_root() -> str:
    """CHANGE3D_REFERENCE_ROOT, else /root/reference (authoring container), else the byte-identical copy that
    `oracle/build_ref.py` staged under oracle/_ref/ (the only one that exists on the GPU box)."""
    env = os.environ.get("CHANGE3D_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/model/trainer.py"):
        return "/root/reference"
    return os.path.join(_HERE, "_ref")


REFERENCE_ROOT = _resolve_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "trainer.py"))


def load():
    """Returns the reference's `model.trainer` module (with Trainer, Encoder)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import model.trainer as ref_trainer  # noqa: E402  (the reference's file, unchanged)
    return ref_trainer


def make_args(task: str, H: int, W: int, num_class: int) -> argparse.Namespace:
    """The constructor namespace the reference scripts pass (scripts/train_BCD.py:410-456 etc.)."""
    P = {"bcd": 1, "bda": 2, "scd": 3}[task]
    dataset = {"bcd": "LEVIR-CD", "bda": "xBD", "scd": "SECOND"}[task]
    return argparse.Namespace(num_perception_frame=P, num_class=num_class, in_height=H, in_width=W,
                              dataset=dataset, pretrained="/nonexistent/X3D_L.pyth")


def build_trainer(task: str, H: int, W: int, num_class: int, state_dict=None):
    import contextlib
    import io
    ref = load()
    with contextlib.redirect_stdout(io.StringIO()):   # the swallowed pretrained-load failure prints
        t = ref.Trainer(make_args(task, H, W, num_class))
    if state_dict is not None:
        t.load_state_dict(state_dict, strict=True)
    return t.float()
