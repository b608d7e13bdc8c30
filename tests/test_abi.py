"""CPU suite: the drop-in boundary.  The product .so must load without a GPU, export every
symbol include/patolette_b200.h declares, and the Python surface must mirror the reference's."""
import ctypes as C
import inspect
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def test_library_builds_and_exports_every_declared_symbol():
    from patolette_b200 import build, _lib
    path = build.build()  # no-op when up to date; nvcc cross-compiles without a GPU
    assert os.path.exists(path)
    header = open(os.path.join(ROOT, "include", "patolette_b200.h")).read()
    declared = set(re.findall(r"^PB200_API[^;(]*?\b(\w+)\s*\(", header, flags=re.M))
    assert {"patolette", "get_patolette_exit_code_info_message", "patolette_create_default_options"} <= declared
    assert declared == set(_lib.SYMBOLS), "ctypes table and header drifted apart"
    nm = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in nm.splitlines() if " T " in ln}
    assert declared <= exported, f"missing exports: {declared - exported}"
    lib = _lib.load()
    for name in declared:
        assert getattr(lib, name) is not None


def test_options_struct_layout_matches_reference_header():
    """lib/include/patolette.h:13-20 on x86-64: offsets 0,1,4,8,16,24, sizeof 32."""
    from patolette_b200._lib import QuantizationOptions as Q
    assert C.sizeof(Q) == 32
    assert [getattr(Q, f).offset for f, _ in Q._fields_] == [0, 1, 4, 8, 16, 24]


def test_default_options_and_messages():
    from patolette_b200 import _lib
    lib = _lib.load()
    o = lib.patolette_create_default_options().contents  # patolette.c:107-119
    assert (o.dither, o.palette_only, o.color_space, o.kmeans_niter, o.kmeans_max_samples, o.verbose) == \
        (True, False, 2, 32, 512 ** 2, False)
    msg = lib.get_patolette_exit_code_info_message
    assert msg(0) == b"Quantization successful."
    assert msg(-1) == b"Internal quantization error."
    assert msg(-2) == b"Image dimensions should be greater than 0."
    assert msg(-3) == b"Palette size should be greater than 0."
    assert msg(-4) == b"Image dimensions are too big."


def test_argument_validation_happens_before_any_cuda_call():
    """patolette.c:61-95 order: dims, palette size, > 40000^2 - all answerable without a GPU."""
    import patolette_b200 as pb
    assert pb.quantize(0, 4, np.zeros((0, 3)), 4, tile_size=0) == (False, None, None, "Image dimensions should be greater than 0.")
    assert pb.quantize(2, 2, np.zeros((4, 3)), 0, tile_size=0)[3] == "Palette size should be greater than 0."


def test_python_surface_mirrors_reference():
    """src/patolette/patolette.pyx:332-344 positional names and defaults."""
    import patolette_b200 as pb
    sig = inspect.signature(pb.quantize)
    names = list(sig.parameters)
    assert names[:11] == ["width", "height", "colors", "palette_size", "dither", "palette_only", "color_space",
                          "tile_size", "kmeans_niter", "kmeans_max_samples", "verbose"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["dither"], d["palette_only"], d["color_space"], d["tile_size"], d["kmeans_niter"],
            d["kmeans_max_samples"], d["verbose"]) == (True, False, 2, 512, 32, 512 ** 2, False)
    assert (pb.ColorSpace_sRGB, pb.ColorSpace_CIELuv, pb.ColorSpace_ICtCp) == (0, 1, 2)
    # python-side validation messages, patolette.pyx:328-330 / :351-373
    assert pb.quantize(2, 2, np.zeros((4, 4)), 2, tile_size=0)[3].startswith("Expected colors to be in sRGB[0, 1] space")
    assert pb.quantize(2, 3, np.zeros((4, 3)), 2, tile_size=0)[3] == "The number of colors doesn't match the supplied width and height."
    assert pb.quantize(2, 2, np.zeros((4, 3)), 2, tile_size=-1)[3] == "tile_size parameter expected to be in the range [0, inf]"


def test_no_cpu_fallback_without_gpu():
    """On a box without a GPU the call must fail loudly (exit code -5), never compute on the host."""
    from patolette_b200 import _lib
    import patolette_b200 as pb
    if _lib.load().patolette_b200_device_count() >= 1:
        pytest.skip("a GPU is present")
    ok, pal, pmap, msg = pb.quantize(4, 4, np.random.rand(16, 3), 4, tile_size=0)
    assert not ok and pal is None and pmap is None and "CUDA" in msg


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under patolette_b200/ may mention it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "patolette_b200")):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle/" not in text and "from oracle" not in text and "import oracle" not in text, f
                assert "libpatolette_oracle" not in text and "libpatolette_ref" not in text, f


def test_pow_restatement_matches_libm_on_host(tmp_path):
    """pow_glibc.h is shared by the CUDA kernels and this host build: 33 M inputs, bit for bit."""
    exe = tmp_path / "pow_check"
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", f"-I{ROOT}/patolette_b200/csrc",
                    f"{ROOT}/tests/native/pow_host_check.c", "-o", str(exe), "-lm"], check=True)
    r = subprocess.run([str(exe), "1000000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "mismatches=0" in r.stdout


def test_route_and_sharding_knobs_validate_their_arguments():
    """patolette_b200_set_option / patolette_b200_set_sharding are pure host state: usable without a GPU."""
    import ctypes as C
    from patolette_b200 import _lib
    import patolette_b200 as pb
    lib = _lib.load()
    assert lib.patolette_b200_set_option(b"no_such_option", 1) == -1
    for name in (b"dump_cap", b"overlap", b"nn_grid", b"dither_grid"):
        assert lib.patolette_b200_set_option(name, 1) == 0
    lib.patolette_b200_set_option(b"dump_cap", -1)
    lib.patolette_b200_set_option(b"overlap", -1)
    assert lib.patolette_b200_set_sharding(0, 0, None, None) == -1      # world < 1
    assert lib.patolette_b200_set_sharding(2, 2, None, None) == -1      # rank out of range
    assert lib.patolette_b200_set_sharding(0, 2, None, None) == -1      # world > 1 needs a callback
    assert lib.patolette_b200_set_sharding(0, 1, None, None) == 0
    with pytest.raises(ValueError):
        pb.set_sharding(0, 2)                                           # no allgather given
    pb.set_sharding(1, 2, lambda send: send * 2)
    pb.set_sharding(0, 1)                                               # back to a single rank


def test_save_png_round_trips(tmp_path):
    """N4: the indexed-PNG writer (host code, no GPU): decode the file by hand and compare."""
    import struct, zlib
    import numpy as np
    import patolette_b200 as pb
    rng = np.random.default_rng(5)
    w, h, K = 37, 11, 64
    pal = np.full((K, 3), -1.0, order="F")
    pal[:40] = rng.random((40, 3))
    pmap = rng.integers(0, 40, w * h).astype(np.uintp)
    path = tmp_path / "q.png"
    size = pb.save_png(str(path), w, h, pal, pmap)
    blob = path.read_bytes()
    assert len(blob) == size and blob[:8] == b"\x89PNG\r\n\x1a\n"
    at, chunks = 8, {}
    while at < len(blob):
        n, tag = struct.unpack(">I4s", blob[at:at + 8])
        data = blob[at + 8:at + 8 + n]
        assert struct.unpack(">I", blob[at + 8 + n:at + 12 + n])[0] == zlib.crc32(tag + data) & 0xffffffff
        chunks[tag] = data
        at += 12 + n
    assert struct.unpack(">IIBBBBB", chunks[b"IHDR"]) == (w, h, 8, 3, 0, 0, 0)
    assert chunks[b"PLTE"] == (pal[:40] * 255).astype(np.uint8).tobytes()
    rows = np.frombuffer(zlib.decompress(chunks[b"IDAT"]), dtype=np.uint8).reshape(h, w + 1)
    assert (rows[:, 0] == 0).all() and np.array_equal(rows[:, 1:].reshape(-1), pmap.astype(np.uint8))
    import pytest
    with pytest.raises(ValueError):
        pb.save_png(str(path), w, h, np.zeros((300, 3)), pmap)


def test_save_gif_round_trips(tmp_path):
    """N4: the GIF writer (fixed-width LZW stream) decodes to the same indices and palette in an independent decoder."""
    Image = pytest.importorskip("PIL.Image")
    import patolette_b200 as pb
    rng = np.random.default_rng(0)
    for (w, h, K, used) in [(37, 29, 256, 256), (64, 64, 16, 11), (5, 3, 4, 2), (254, 3, 256, 200), (508, 2, 256, 256), (300, 200, 2, 2)]:
        pal = np.full((K, 3), -1.0)
        pal[:used] = rng.random((used, 3))
        idx = rng.integers(0, used, w * h)
        path = str(tmp_path / "t.gif")
        assert pb.save_gif(path, w, h, pal, idx) == os.path.getsize(path)
        im = Image.open(path)
        a = np.array(im)
        assert im.mode == "P" and a.shape == (h, w)
        assert np.array_equal(a.reshape(-1), idx)
        assert np.array_equal(np.array(im.getpalette()[:3 * used]).reshape(used, 3), (pal[:used] * 255).astype(np.uint8))
    with pytest.raises(ValueError):
        pb.save_gif(str(tmp_path / "x.gif"), 2, 2, np.zeros((300, 3)), np.zeros(4, dtype=int))
