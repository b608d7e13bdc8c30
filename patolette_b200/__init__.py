"""patolette_b200 - B200-native drop-in for the pixel-array hot path of big-nacho/patolette.

Public surface = the reference's Python module (src/patolette/__init__.py:1-10):
``quantize``, ``ColorSpace_sRGB``, ``ColorSpace_CIELuv``, ``ColorSpace_ICtCp``.
Everything runs through the C ABI of libpatolette_b200.so; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

__version__ = "0.1.0"

# reference: src/patolette/patolette.pyx:324-326
ColorSpace_sRGB = 0
ColorSpace_CIELuv = 1
ColorSpace_ICtCp = 2

# reference: src/patolette/patolette.pyx:328-330
color_mismatch = "The number of colors doesn't match the supplied width and height."
bad_channel_count = "Expected colors to be in sRGB[0, 1] space. Channel count mismatch: {} found."
bad_tile_size = "tile_size parameter expected to be in the range [0, inf]"

__all__ = ["__doc__", "__version__", "quantize", "ColorSpace_sRGB", "ColorSpace_CIELuv", "ColorSpace_ICtCp",
           "quantize_u8", "saliency_weights", "saliency_mbd", "save_png", "save_gif", "init_sharding", "shard_range", "quantize_sharded", "sharding_description",
           "set_sharding", "torch_allgather"]


def quantize(width, height, colors, palette_size, dither=True, palette_only=False,
             color_space=ColorSpace_ICtCp, tile_size=512, kmeans_niter=32,
             kmeans_max_samples=512 ** 2, verbose=False, *, weights=None):
    """Same contract as the reference ``quantize`` (src/patolette/patolette.pyx:332-466):
    returns ``(success, palette[K,3] F-order f64 | None, palette_map[N] uintp | None, message)``.

    ``tile_size > 0`` (the default, 512) weights the pixels by saliency as the reference's wrapper does
    (``get_weights``, patolette.pyx:203-313): here the weights are computed on the GPU (pb_saliency.cu) from the sRGB
    input and never leave the device.  The minimum-barrier distance map is the reference's bit for bit; the Lab /
    Mahalanobis / sigmoid chain agrees with numpy + scikit-image to ~1e-13 relative, not to the last bit, so a
    weighted palette can differ from the reference's in its last digits (an unweighted one, ``tile_size=0``, never
    does).  Where the reference's wrapper raises (a side <= 3, border strips that do not fit the image) this raises
    ``ValueError``.

    Extension: ``weights`` (keyword-only) - per-pixel f64 weights >= 1 handed straight to the C ABI's ``weights``
    argument (lib/include/patolette.h:26); ``tile_size`` is then ignored.
    """
    colors = np.asarray(colors)
    if colors.ndim != 2:
        raise ValueError("colors must be a 2-D array")
    color_count, channel_count = colors.shape
    if channel_count != 3:
        return (False, None, None, bad_channel_count.format(channel_count))
    if color_count != width * height:
        return (False, None, None, color_mismatch)
    if tile_size < 0:
        return (False, None, None, bad_tile_size)
    lib = _lib.load()
    # The reference copies to Fortran order on the host (patolette.pyx:388-391).  A C-contiguous f64
    # array is instead handed over as is and de-interleaved on the GPU (same values, no host pass).
    interleaved = colors.dtype == np.float64 and colors.flags.c_contiguous and not colors.flags.f_contiguous
    data = colors if interleaved else np.asfortranarray(colors, dtype=np.float64)
    palette = np.zeros((palette_size, 3), dtype=np.float64, order="F")
    pmap = None if palette_only else np.zeros(width * height, dtype=np.uintp)
    w = None
    if weights is not None:
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if w.shape != (color_count,):
            raise ValueError("weights must have one entry per pixel")
    opts = _lib.QuantizationOptions(bool(dither), bool(palette_only), int(color_space), int(kmeans_niter),
                                    int(kmeans_max_samples), bool(verbose))
    code = C.c_int(0)
    if w is None and tile_size > 0:  # patolette.pyx:411-415
        if verbose:
            print("patolette ======== Generating saliency map")
        lib.patolette_b200_quantize(width, height, data.ctypes.data if color_count else None, 1 if interleaved else 0,
                                    float(tile_size), palette_size, C.byref(opts),
                                    palette.ctypes.data if palette_size else None,
                                    None if pmap is None else pmap.ctypes.data, 8, 0, C.byref(code))
        if code.value == -7:
            raise ValueError(lib.get_patolette_exit_code_info_message(-7).decode("utf-8"))
    else:
        entry = lib.patolette_b200_interleaved if interleaved else lib.patolette
        entry(width, height, data.ctypes.data if color_count else None,
              None if w is None else w.ctypes.data, palette_size, C.byref(opts),
              palette.ctypes.data if palette_size else None,
              None if pmap is None else pmap.ctypes.data, C.byref(code))
    global _shard_error
    if _shard_error is not None:  # the all-gather callback failed: surface it instead of exit code -1
        err, _shard_error = _shard_error, None
        raise err
    success = code.value == 0
    message = lib.get_patolette_exit_code_info_message(code.value).decode("utf-8")
    if not success:
        return (success, None, None, message)
    if palette_only:
        return (success, palette, None, message)
    return (success, palette, pmap, message)


def quantize_u8(width, height, rgb, palette_size, dither=True, palette_only=False, color_space=ColorSpace_ICtCp,
                kmeans_niter=32, kmeans_max_samples=512 ** 2, verbose=False, *, weights=None, tile_size=0):
    """N1 (extension): quantise an 8-bit image without the host-side f64 conversion of README.md:150-158.

    ``rgb``: uint8 array [N, 3] (or [H, W, 3]).  The division by 255 happens on the device in f64 (the same IEEE
    operation), so the result equals ``quantize(width, height, rgb / 255.0, ..., tile_size=tile_size)``; the map comes
    back as uint8 (palette_size <= 256) or uint16 instead of uintp.  ``tile_size > 0`` adds the saliency weights
    (see :func:`quantize`); note the default here is 0.  Returns (success, palette, palette_map, message)."""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3)
    if rgb.shape[0] != width * height:
        return (False, None, None, color_mismatch)
    if palette_size > 65536:
        raise ValueError("quantize_u8 returns 8- or 16-bit indices: palette_size <= 65536")
    lib = _lib.load()
    map_dtype = np.uint8 if palette_size <= 256 else np.uint16
    palette = np.zeros((palette_size, 3), dtype=np.float64, order="F")
    pmap = None if palette_only else np.zeros(width * height, dtype=map_dtype)
    w = None
    if weights is not None:
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if w.shape != (width * height,):
            raise ValueError("weights must have one entry per pixel")
    opts = _lib.QuantizationOptions(bool(dither), bool(palette_only), int(color_space), int(kmeans_niter),
                                    int(kmeans_max_samples), bool(verbose))
    code = C.c_int(0)
    if tile_size < 0:
        return (False, None, None, bad_tile_size)
    if w is None and tile_size > 0:
        lib.patolette_b200_quantize(width, height, rgb.ctypes.data if rgb.size else None, 2, float(tile_size), palette_size,
                                    C.byref(opts), palette.ctypes.data if palette_size else None,
                                    None if pmap is None else pmap.ctypes.data, np.dtype(map_dtype).itemsize, 0,
                                    C.byref(code))
        if code.value == -7:
            raise ValueError(lib.get_patolette_exit_code_info_message(-7).decode("utf-8"))
    else:
        lib.patolette_b200_u8(width, height, rgb.ctypes.data if rgb.size else None, None if w is None else w.ctypes.data,
                              palette_size, C.byref(opts), palette.ctypes.data if palette_size else None,
                              None if pmap is None else pmap.ctypes.data, np.dtype(map_dtype).itemsize, 0, C.byref(code))
    success = code.value == 0
    message = lib.get_patolette_exit_code_info_message(code.value).decode("utf-8")
    if not success:
        return (success, None, None, message)
    return (success, palette, None if palette_only else pmap, message)


def _stage_error(lib, rc: int, what: str):
    if rc == -7:
        raise ValueError(lib.get_patolette_exit_code_info_message(-7).decode("utf-8"))
    if rc != 0:
        raise RuntimeError(f"patolette_b200: {what} failed ({rc})")


def saliency_weights(width, height, colors, tile_size=512):
    """N3 stage (extension): the per-pixel weights ``quantize(..., tile_size)`` uses - ``get_weights(img, tile_size)``
    of the reference's wrapper (patolette.pyx:203-313) on the GPU.  ``colors``: [N, 3] sRGB in [0, 1], pixel
    p = row * width + col.  Returns N float64 weights (each >= 1)."""
    colors = np.asarray(colors)
    if colors.ndim != 2 or colors.shape[1] != 3 or colors.shape[0] != width * height:
        raise ValueError(color_mismatch)
    if not tile_size > 0:
        raise ValueError(bad_tile_size)
    lib = _lib.load()
    data = np.asfortranarray(colors, dtype=np.float64)
    out = np.empty(width * height, dtype=np.float64)
    _stage_error(lib, lib.patolette_b200_saliency_weights(width, height, data.ctypes.data, float(tile_size),
                                                          out.ctypes.data, 0), "saliency_weights")
    return out


def saliency_mbd(width, height, colors):
    """N3 stage (extension): the minimum-barrier distance map ``mbd(mean(img, axis=2).astype(float32), 3)``
    (patolette.pyx:153-201) as a [height, width] float32 array - bit for bit the reference's."""
    colors = np.asarray(colors)
    if colors.ndim != 2 or colors.shape[1] != 3 or colors.shape[0] != width * height:
        raise ValueError(color_mismatch)
    lib = _lib.load()
    data = np.asfortranarray(colors, dtype=np.float64)
    out = np.empty((height, width), dtype=np.float32)
    _stage_error(lib, lib.patolette_b200_saliency_mbd(width, height, data.ctypes.data, out.ctypes.data, 0), "saliency_mbd")
    return out


def save_png(path, width, height, palette, palette_map, compress_level=6):
    """N4 (extension): write ``(palette, palette_map)`` as an indexed PNG - the step that follows ``quantize()`` in the
    reference's README (README.md:186-191 reassembles the image with PIL).  Colour type 3, bit depth 8 (palettes up
    to 256 entries; unused rows - marked -1 by the reference, patolette.c:328-330 - are dropped from PLTE), palette
    entries rounded as the README does (``(palette * 255).astype(uint8)`` truncates; so does this).  Pure host code
    on the standard library (zlib + struct); returns the number of bytes written."""
    import struct
    import zlib
    pal = np.asarray(palette, dtype=np.float64)
    idx = np.asarray(palette_map).reshape(-1)
    if pal.ndim != 2 or pal.shape[1] != 3:
        raise ValueError("palette must be K x 3")
    if idx.size != width * height:
        raise ValueError(color_mismatch)
    used = int((pal[:, 0] >= 0).sum()) if pal.size else 0  # trailing rows of -1 are unused slots
    if used > 256:
        raise ValueError("an indexed PNG holds at most 256 palette entries")
    if idx.size and int(idx.max()) >= max(used, 1):
        raise ValueError("palette_map refers to an unused palette row")
    plte = (np.clip(pal[:used], 0.0, 1.0) * 255).astype(np.uint8).tobytes()
    rows = np.zeros((height, width + 1), dtype=np.uint8)  # filter byte 0 (None) + one index per pixel
    rows[:, 1:] = idx.astype(np.uint8).reshape(height, width)

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)

    blob = (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", width, height, 8, 3, 0, 0, 0))
            + chunk(b"PLTE", plte) + chunk(b"IDAT", zlib.compress(rows.tobytes(), compress_level)) + chunk(b"IEND", b""))
    with open(path, "wb") as f:
        f.write(blob)
    return len(blob)


def save_gif(path, width, height, palette, palette_map):
    """N4 (extension): write ``(palette, palette_map)`` as a GIF89a (global colour table, one image) - palettes up to
    256 entries, entries rounded as the reference's README does (``(palette * 255).astype(uint8)``, README.md:186-191).
    The LZW stream is of the "uncompressed" kind (fixed code width, a clear code before the decoder's table would
    grow): every decoder reads it, the packing is a handful of numpy operations instead of a per-pixel Python loop,
    and the file is (bits + 1) / 8 bytes per pixel.  Pure host code; returns the number of bytes written."""
    import struct
    pal = np.asarray(palette, dtype=np.float64)
    idx = np.asarray(palette_map).reshape(-1)
    if pal.ndim != 2 or pal.shape[1] != 3:
        raise ValueError("palette must be K x 3")
    if idx.size != width * height:
        raise ValueError(color_mismatch)
    if not (0 < width < 65536 and 0 < height < 65536):
        raise ValueError("a GIF is at most 65535 pixels wide and high")
    used = int((pal[:, 0] >= 0).sum()) if pal.size else 0  # trailing rows of -1 are unused slots
    if used > 256:
        raise ValueError("a GIF holds at most 256 palette entries")
    if idx.size and int(idx.max()) >= max(used, 1):
        raise ValueError("palette_map refers to an unused palette row")
    m = max(2, int(np.ceil(np.log2(max(used, 2)))))  # LZW minimum code size = bits per index
    table = np.zeros((1 << m, 3), dtype=np.uint8)
    table[:used] = (np.clip(pal[:used], 0.0, 1.0) * 255).astype(np.uint8)
    clear, eoi, wbits = 1 << m, (1 << m) + 1, m + 1
    run = (1 << m) - 2  # data codes between two clear codes: the decoder's table never reaches 2^(m+1) entries
    n = idx.size
    nruns = (n + run - 1) // run
    body = np.empty((nruns, run + 1), dtype=np.uint16)  # [clear, run codes] per row
    body[:, 0] = clear
    flat = np.full(nruns * run, eoi, dtype=np.uint16)
    flat[:n] = idx.astype(np.uint16)
    body[:, 1:] = flat.reshape(nruns, run)
    # (every run contributes its clear code and its pixels; the last one may be partial)
    stream = np.concatenate([body.reshape(-1)[:n + nruns], np.array([eoi], dtype=np.uint16)])
    bits = ((stream[:, None] >> np.arange(wbits, dtype=np.uint16)) & 1).astype(np.uint8).reshape(-1)
    data = np.packbits(bits, bitorder="little")
    nblk = (data.size + 254) // 255
    padded = np.zeros(nblk * 255, dtype=np.uint8)
    padded[:data.size] = data
    blocks = np.empty((nblk, 256), dtype=np.uint8)
    blocks[:, 0] = 255
    blocks[:, 1:] = padded.reshape(nblk, 255)
    tail = data.size - (nblk - 1) * 255
    blocks[-1, 0] = tail
    image_data = blocks.reshape(-1)[:(nblk - 1) * 256 + 1 + tail].tobytes() + b"\x00"
    blob = (b"GIF89a" + struct.pack("<HHBBB", width, height, 0xF0 | (m - 1), 0, 0) + table.tobytes()
            + b"," + struct.pack("<HHHHB", 0, 0, width, height, 0) + bytes([m]) + image_data + b";")
    with open(path, "wb") as f:
        f.write(blob)
    return len(blob)


def init_sharding(dist=None, group=None) -> tuple[int, int]:
    """Image-sharded multi-GPU runs (extension): create the library's NCCL communicator over the
    ranks of a torch.distributed process group (one process per GPU; call after ``torch.cuda.set_device`` /
    ``patolette_b200_set_device``).  The 128-byte NCCL unique id travels through ``dist.broadcast``; torch is
    plumbing only.  Returns (rank, world)."""
    import torch
    if dist is None:
        import torch.distributed as dist
    lib = _lib.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ident = C.create_string_buffer(128)
    if rank == 0 and lib.patolette_b200_comm_unique_id(ident) != 0:
        raise RuntimeError("patolette_b200: NCCL is not available (libnccl.so.2)")
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.frombuffer(bytearray(ident.raw), dtype=torch.uint8).to(dev)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    raw = bytes(t.cpu().numpy().tobytes())
    if lib.patolette_b200_comm_init(rank, world, raw) != 0:
        raise RuntimeError("patolette_b200: ncclCommInitRank failed")
    return rank, world


def shard_range(n_pixels: int, rank: int, world: int) -> tuple[int, int]:
    """(first, count): the pixels rank ``rank`` of ``world`` brings to :func:`quantize_sharded`."""
    first, count = C.c_size_t(0), C.c_size_t(0)
    if _lib.load().patolette_b200_shard_range(n_pixels, rank, world, C.byref(first), C.byref(count)) != 0:
        raise ValueError("bad rank / world")
    return first.value, count.value


def sharding_description() -> str:
    lib = _lib.load()
    r, w, v = C.c_int(0), C.c_int(1), C.c_int(0)
    lib.patolette_b200_comm_info(C.byref(r), C.byref(w), C.byref(v))
    return (f"image-sharded x{w.value}: one image, pixel slices in / map slices out; colour planes all-gathered over NVLink "
            f"(NCCL {v.value}), split loop sharded by cluster with a device-side all-gather per batch, GQ / KMeans sample / "
            f"dither walk replicated")


def quantize_sharded(width, height, colors_slice, palette_size, dither=True, palette_only=False,
                     color_space=ColorSpace_ICtCp, kmeans_niter=32, kmeans_max_samples=512 ** 2, verbose=False, *,
                     weights_slice=None):
    """Collective over the ranks of :func:`init_sharding`: every rank passes ITS pixels (``shard_range``) as an
    [count, 3] f64 array and gets back ``(success, palette, palette_map_of_its_pixels, message)``.  Bit-identical to
    ``quantize`` on the whole image for every world size."""
    lib = _lib.load()
    r, w = C.c_int(0), C.c_int(1)
    if not lib.patolette_b200_comm_info(C.byref(r), C.byref(w), None):
        raise RuntimeError("call init_sharding() first")
    first, count = shard_range(width * height, r.value, w.value)
    data = np.asfortranarray(colors_slice, dtype=np.float64)
    if data.shape != (count, 3):
        raise ValueError(f"rank {r.value} of {w.value} must pass pixels [{first}, {first + count}): shape ({count}, 3)")
    ws = None
    if weights_slice is not None:
        ws = np.ascontiguousarray(weights_slice, dtype=np.float64)
        if ws.shape != (count,):
            raise ValueError("weights_slice must have one entry per pixel of the slice")
    palette = np.zeros((palette_size, 3), dtype=np.float64, order="F")
    pmap = None if palette_only else np.zeros(count, dtype=np.uintp)
    opts = _lib.QuantizationOptions(bool(dither), bool(palette_only), int(color_space), int(kmeans_niter),
                                    int(kmeans_max_samples), bool(verbose))
    code = C.c_int(0)
    lib.patolette_b200_sharded(width, height, data.ctypes.data if count else None, None if ws is None else ws.ctypes.data,
                               palette_size, C.byref(opts), palette.ctypes.data,
                               None if pmap is None or not count else pmap.ctypes.data, 0, C.byref(code))
    success = code.value == 0
    message = lib.get_patolette_exit_code_info_message(code.value).decode("utf-8")
    if not success:
        return (success, None, None, message)
    return (success, palette, pmap, message)


_shard_cb = None  # keeps the ctypes callback alive
_shard_error = None  # exception raised inside the all-gather callback of the running call


def set_sharding(rank: int, world: int, allgather=None) -> None:
    """Chain-sharded multi-GPU runs (extension).  Every rank - one process per GPU - calls
    ``quantize`` with the SAME image at the same time; the ordered sums are split over the ranks by chain and
    the per-cluster moment rows are exchanged through ``allgather(send: bytes) -> bytes`` (the concatenation
    of every rank's ``send``, rank-major; e.g. built on ``torch.distributed.all_gather``).  Results are
    bit-identical for every world size.  ``world == 1`` switches sharding off."""
    global _shard_cb
    lib = _lib.load()
    if world <= 1:
        _shard_cb = None
        if lib.patolette_b200_set_sharding(0, 1, None, None) != 0:
            raise ValueError("bad sharding arguments")
        return
    if allgather is None:
        raise ValueError("world > 1 needs an allgather callable")

    @C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
    def cb(send, recv, nbytes, _user):
        # ctypes swallows exceptions raised inside a callback: catch, stash, and tell the C side to abort
        # the call (exit code -1); quantize() re-raises the stashed exception
        global _shard_error
        try:
            out = allgather(C.string_at(send, nbytes))
            if len(out) != nbytes * world:
                raise RuntimeError("allgather returned %d bytes, expected %d" % (len(out), nbytes * world))
            C.memmove(recv, out, len(out))
            return 0
        except BaseException as e:  # noqa: BLE001 - must not propagate into C
            _shard_error = e
            return 1

    if lib.patolette_b200_set_sharding(int(rank), int(world), C.cast(cb, C.c_void_p), None) != 0:
        raise ValueError("bad sharding arguments")
    _shard_cb = cb


def torch_allgather(group=None):
    """An ``allgather`` for :func:`set_sharding` on top of torch.distributed (host tensors: use a gloo group, or
    the default group when it is gloo)."""
    import torch
    import torch.distributed as dist

    def allgather(send: bytes) -> bytes:
        t = torch.frombuffer(bytearray(send), dtype=torch.uint8)
        outs = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
        dist.all_gather(outs, t, group=group)
        return b"".join(o.numpy().tobytes() for o in outs)

    return allgather


def last_timings() -> dict:
    """Stage timings (ms) of the last quantize() in this process, for bench.py."""
    out = (C.c_double * 10)()
    _lib.load().patolette_b200_last_timings(out)
    keys = ["total", "h2d", "color", "gq", "lq", "kmeans", "nearest", "dither", "d2h", "launches"]
    t = dict(zip(keys, list(out)))
    t["saliency"] = float(_lib.load().patolette_b200_last_saliency_ms())  # part of "color"
    return t
