"""Execute the reference's own source on CPU (container only — needs /root/reference).

TEST INFRASTRUCTURE.  `load_gan_utils()` imports `/root/reference/gan_utils.py` UNMODIFIED with
`oracle/tf_shim` standing in for TensorFlow; `load_kernel_smoothing()` execs the class body at
`data_utils.py:478-586` (the whole module cannot be imported: cv2/matplotlib/IPython/tf.keras
imports at the top).  No reference source is copied into this repo: the text is read from
`/root/reference` at call time.
"""
import importlib.util
import os
import sys

REFERENCE_ROOT = os.environ.get("KCCOT_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf_shim")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "gan_utils.py"))


def _with_shim():
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    import tensorflow  # noqa: F401  (the shim)
    return sys.modules["tensorflow"]


def load_gan_utils():
    """The reference `gan_utils` module, executing on torch CPU tensors."""
    if not reference_available():
        raise FileNotFoundError(f"{REFERENCE_ROOT}/gan_utils.py (only present in the build container)")
    _with_shim()
    spec = importlib.util.spec_from_file_location(
        "_kccot_reference_gan_utils", os.path.join(REFERENCE_ROOT, "gan_utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_kernel_smoothing():
    """The reference `KernelSmoothing` class (data_utils.py:478-586), executing on torch CPU."""
    if not reference_available():
        raise FileNotFoundError(f"{REFERENCE_ROOT}/data_utils.py (only present in the build container)")
    tf = _with_shim()
    import numpy as np
    with open(os.path.join(REFERENCE_ROOT, "data_utils.py")) as f:
        lines = f.read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("class KernelSmoothing"))
    end = next(i for i in range(start + 1, len(lines))
               if lines[i].startswith("class ") or lines[i].startswith("def "))
    ns = {"tf": tf, "np": np}
    exec(compile("\n".join(lines[start:end]), os.path.join(REFERENCE_ROOT, "data_utils.py"), "exec"), ns)
    return ns["KernelSmoothing"]
