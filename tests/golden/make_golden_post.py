"""Golden vectors for the caller-side colour corrections, made by executing the REFERENCE's own
``correct_bayer_channels`` and ``auto_correct_rb``.  test.py cannot be imported here (its top-level imports need skimage
and imageio), so the two function definitions are cut out of its syntax tree and executed as they are.

    python tests/golden/make_golden_post.py        # needs /root/reference; writes tests/golden/post.npz
"""
import ast
import os

import numpy as np

REF = "/root/reference/test.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "post.npz")


def reference_functions():
    tree = ast.parse(open(REF).read())
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("correct_bayer_channels", "auto_correct_rb")]
    assert len(wanted) == 2
    ns = {"np": np}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), REF, "exec"), ns)
    return ns["correct_bayer_channels"], ns["auto_correct_rb"]


def inputs():
    """name -> uint8 HWC image: red-heavy, blue-heavy, equal means, tiny."""
    rng = np.random.default_rng(2024)
    a = rng.integers(0, 256, size=(12, 20, 3), dtype=np.uint8)
    red = a.copy(); red[..., 0] = np.maximum(red[..., 0], 128)
    blue = a.copy(); blue[..., 2] = np.maximum(blue[..., 2], 160); blue[..., 0] //= 2
    equal = a.copy(); equal[..., 2] = equal[..., 0]
    green = a.copy(); green[..., 1] = 255; green[..., 0] //= 3           # G stronger than R: matters after GBRG / GRBG
    return {"red": red, "blue": blue, "equal": equal, "green": green, "tiny": a[:1, :2].copy()}


if __name__ == "__main__":
    cbc, arb = reference_functions()
    out = {}
    for name, img in inputs().items():
        out["in_" + name] = img
        for pat in ("RGGB", "BGGR", "GBRG", "GRBG", "rggb", "XXXX"):
            out[f"out_{name}_{pat}"] = np.ascontiguousarray(arb(cbc(img.copy(), pat)))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, len(out), "arrays")
