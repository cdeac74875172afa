"""Import the unmodified reference (`gym_socks`) for measurement and tests (test infrastructure).

Search order: /root/reference (authoring container), then oracle/_ref/gym_socks_ref.zip (the verbatim archive made by
oracle/vendor_ref.py, which is what exists on the GPU box; imported in place with zipimport).  `gym` is not installed in this image, so the test-only
shim in oracle/gym_shim is put on the path first."""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))


ARCHIVE = os.path.join(HERE, "_ref", "gym_socks_ref.zip")


def reference_root():
    """Path to put on sys.path: the reference checkout, or the vendored archive."""
    if os.path.isfile(os.path.join("/root/reference", "gym_socks", "algorithms", "reach", "kernel_sr.py")):
        return "/root/reference"
    if os.path.isfile(ARCHIVE):
        return ARCHIVE
    return None


def load():
    """-> dict of the reference's hot-path classes/functions, or raises ImportError with the reason."""
    root = reference_root()
    if root is None:
        raise ImportError("the reference is neither at /root/reference nor vendored as oracle/_ref/gym_socks_ref.zip "
                          "(run `python oracle/vendor_ref.py` where /root/reference exists)")
    sys.dont_write_bytecode = True  # /root/reference is read-only
    for p in (root, os.path.join(HERE, "gym_shim")):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    import gym  # noqa: F401  (the shim)
    from gym_socks.algorithms.reach.kernel_sr import KernelSR
    from gym_socks.kernel.metrics import rbf_kernel, regularized_inverse

    return {"root": root, "gym": gym, "KernelSR": KernelSR, "rbf_kernel": rbf_kernel,
            "regularized_inverse": regularized_inverse}
