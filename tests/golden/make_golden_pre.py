"""Golden vectors for the RAW normalisation, made by executing the REFERENCE's own statements.

``RawFomer_WFB_FFAB/load_dataset.py`` cannot be imported here (rawpy / imageio are absent), so the two assignment
statements of ``load_data_SID.__getitem__`` that normalise the short exposure (``np.clip(...astype(np.float32), 512,
16383)`` and ``(x - 512) / (16383 - 512 + 1e-6) * ap``, lines 88-89) and the clamp of ``correctdataloader.py`` (line 103,
``np.minimum(img_short, 1.0)``) are cut out of the files' syntax trees and executed as they are on a seeded uint16 frame.

    python tests/golden/make_golden_pre.py        # needs /root/reference; writes tests/golden/pre.npz
"""
import ast
import os

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pre.npz")


def _assignments(path, target, must_contain):
    """The ``target = ...`` statements of a file whose source contains ``must_contain`` (in file order)."""
    src = open(path).read()
    found = []
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name) \
                and node.targets[0].id == target:
            seg = ast.get_source_segment(src, node)
            if any(m in seg for m in must_contain):
                found.append((node.lineno, node))
    return [n for _, n in sorted(found, key=lambda t: t[0])]


def reference_normalise():
    stmts = _assignments(os.path.join(REF, "RawFomer_WFB_FFAB", "load_dataset.py"), "img_short_crop",
                         ("np.clip(img_short_crop.astype", "(img_short_crop - 512)"))
    # the same two statements appear in both dataset classes; keep the first pair (load_data_SID)
    assert len(stmts) >= 2, len(stmts)
    norm = compile(ast.Module(body=stmts[:2], type_ignores=[]), "load_dataset.py", "exec")
    clamp = _assignments(os.path.join(REF, "correctdataloader.py"), "img_short", ("np.minimum(img_short",))
    assert len(clamp) >= 1
    clamp = compile(ast.Module(body=clamp[:1], type_ignores=[]), "correctdataloader.py", "exec")

    def run(raw_u16, ap, do_clamp):
        ns = {"np": np, "img_short_crop": raw_u16, "ap": ap}
        exec(norm, ns)
        x = ns["img_short_crop"]
        if do_clamp:
            ns2 = {"np": np, "img_short": x}
            exec(clamp, ns2)
            x = ns2["img_short"]
        return np.ascontiguousarray(x)

    return run


if __name__ == "__main__":
    run = reference_normalise()
    rng = np.random.default_rng(3)
    raw = rng.integers(0, 16384, size=(2, 16, 24)).astype(np.uint16)
    raw[0, 0, :6] = [0, 511, 512, 513, 16383, 65535]          # below black, at black, at/above white
    out = {"raw": raw}
    for ap in (100, 300):
        for do_clamp in (False, True):
            y = run(raw, ap, do_clamp)
            assert y.dtype == np.float32, y.dtype
            out[f"out_ap{ap}_{'clamp' if do_clamp else 'noclamp'}"] = y
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: (v.shape, str(v.dtype), float(v.max())) for k, v in out.items()})
