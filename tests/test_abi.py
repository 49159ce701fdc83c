"""CPU checks of the C-ABI boundary: the in-tree library builds, loads, exports every symbol the header declares,
and its host-side queries (no kernel launches) behave.  No GPU needed."""
import ctypes
import os
import re

import pytest

import rf_testlib as T


@pytest.fixture(scope="module")
def lib():
    from bayer_low_light_image_enhancement_b200 import _lib, build

    build.build()
    return _lib.load()


def _declared():
    text = open(os.path.join(T.ROOT, "include", "rawformer_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 30
    raw = ctypes.CDLL(lib._name)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/rawformer_b200.h but not exported"


def test_binding_covers_header():
    from bayer_low_light_image_enhancement_b200 import exported_symbols

    assert set(_declared()) == set(exported_symbols())


def test_struct_layout_matches_header(tmp_path):
    """sizeof/offsetof of the ctypes mirrors against the C compiler's view of include/rawformer_b200.h."""
    import subprocess

    from bayer_low_light_image_enhancement_b200._lib import Band, BlockWeights, ModelWeights

    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "rawformer_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %d %d\\n", sizeof(rf_block_weights), sizeof(rf_model_weights),'
        ' offsetof(rf_block_weights, pyr_low_w), offsetof(rf_block_weights, pyr_res_b2),'
        ' offsetof(rf_model_weights, down_w), offsetof(rf_model_weights, rgb_w_host),'
        ' sizeof(rf_band), offsetof(rf_band, comm), offsetof(rf_band, epoch), RF_BAND_MAX_RANKS, RF_BAND_HALO);return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(T.ROOT, "include"), str(src), "-o", str(exe)], check=True)
    c = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    ours = [ctypes.sizeof(BlockWeights), ctypes.sizeof(ModelWeights), BlockWeights.pyr_low_w.offset,
            BlockWeights.pyr_res_b2.offset, ModelWeights.down_w.offset, ModelWeights.rgb_w_host.offset,
            ctypes.sizeof(Band), Band.comm.offset, Band.epoch.offset]
    assert ours == c[:9]
    from bayer_low_light_image_enhancement_b200 import rowtiled

    assert [rowtiled.RF_BAND_MAX_RANKS, rowtiled.RF_BAND_HALO] == c[9:]


def test_host_queries(lib):
    assert lib.rf_version() >= 100
    assert lib.rf_strerror(0) == b"ok"
    assert b"sm_100" in lib.rf_strerror(-3)
    assert lib.rf_kernel_name(0) == b"pack_luma"
    # packed-parameter blob grows with dim, bf16 is smaller than fp32, ML carries the pyramid extras
    s32, s16 = lib.rf_model_packed_bytes(32, 0, 0), lib.rf_model_packed_bytes(32, 1, 0)
    assert s32 > s16 > 2_477_465 * 2
    assert lib.rf_model_packed_bytes(64, 1, 0) > lib.rf_model_packed_bytes(48, 1, 0) > s16
    assert lib.rf_model_packed_bytes(32, 1, 1) > s16
    assert lib.rf_model_packed_bytes(36, 1, 0) == 0  # dim % 8
    # workspace queries are dry runs of the same plan the forward executes
    w_small = lib.rf_rawformer_workspace_bytes(32, 1, 0, 1, 64, 64)
    w_full = lib.rf_rawformer_workspace_bytes(32, 1, 0, 1, 2848, 4256)
    assert 0 < w_small < w_full < 8 << 30
    assert lib.rf_rawformer_workspace_bytes(32, 1, 0, 1, 40, 64) == 0  # H % 16
    assert lib.rf_rawformer_workspace_bytes(32, 1, 0, 2, 64, 64) > w_small
    assert lib.rf_block_workspace_bytes(32, 0, 1, 16, 24, 16, 24) > 0
    assert lib.rf_block_workspace_bytes(33, 0, 1, 16, 24, 16, 24) == 0


def test_argument_errors_without_gpu(lib):
    # null pointers / bad shapes are rejected before any CUDA call
    assert lib.rf_downshuffle(None, None, 1, 1, 4, 4, 2, None) == -2
    k = (ctypes.c_float * 16)()
    assert lib.rf_custom_dwt(ctypes.c_void_p(16), ctypes.c_void_p(16), k, 1, 1, 5, 4, None) == -1
    assert lib.rf_rawformer_forward(None, 32, 0, 0, None, None, 1, 64, 64, None, 0, None) == -2


def test_modules_fail_loudly_on_cpu():
    import torch

    import bayer_low_light_image_enhancement_b200 as rf

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rf.RawFormer(dim=32)(torch.zeros(1, 1, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rf.CustomDWT()(torch.zeros(1, 1, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rf.downshuffle(torch.zeros(1, 1, 4, 4), 2)


def test_state_dict_contract():
    import bayer_low_light_image_enhancement_b200 as rf

    m = rf.RawFormer(model_size="S")
    keys = list(m.state_dict())
    assert len(keys) == 246 and sum(p.numel() for p in m.parameters()) == 2_477_465
    assert "down1.body.0.weight" in keys and "conv_tran1.FLCA.dwt.filt" in keys
    assert "conv_tran7.Transformer.attn.temperature" in keys and "luma_chroma.r_w" in keys
    assert sum(p.numel() for p in rf.RawFormer(model_size="B").parameters()) == 5_516_721
    assert sum(p.numel() for p in rf.RawFormer(model_size="L").parameters()) == 9_756_361
    ml = rf.multilevel.RawFormer(dim=32)
    mk = list(ml.state_dict())
    assert len(mk) == 310 and "down1.0.weight" in mk and "haar.filt" in mk
    assert sum(p.numel() for p in ml.parameters()) == 2_708_710
    assert rf.WaveTransformBlock is rf.Conv_Transformer


def test_streaming_api_surface():
    """FramePipeline / CUDA-graph mode exist on the public surface and refuse to run without a CUDA device."""
    import pytest
    import torch

    import bayer_low_light_image_enhancement_b200 as rf

    m = rf.RawFormer(dim=32)
    assert hasattr(m, "enable_cuda_graphs") and m.enable_cuda_graphs(False) is m
    with pytest.raises(RuntimeError):
        rf.FramePipeline(m)          # parameters on the CPU: there is no CPU path
