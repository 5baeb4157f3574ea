"""The .vdb reader on compressed files: zlib and Blosc(LZ4) blocks, with and without the active
mask, half and full floats.  The shipped assets exercise none of this (compression flag 0x2 only), so the files are
made by tests/vdb_writer.py -- a writer derived from the format the reference's READER expects (read.rs), with a
from-scratch c-blosc 1.x / LZ4 encoder.  Product reader and oracle reader are independent implementations; both must
recover exactly the voxels that were written.  The reference decodes Blosc through the c-blosc C library
(blosc-src 0.2.1, not vendored); the oracle has no such decoder and must say so (WXO_ERR_BLOSC)."""
import numpy as np
import pytest

import oracle_ffi as O
import scenes
import vdb_writer as V
import woxel_b200 as W


def sample_voxels(seed=3):
    rng = np.random.default_rng(seed)
    pts = [rng.integers(0, 8, size=(200, 3)) + base for base in ([0, 0, 0], [8, 0, 0], [120, 120, 120], [128, 0, 0], [-8, -8, -8], [-4096, 16, 24],
                                                                  [4000, 4000, 4000])]
    dense = np.stack(np.meshgrid(np.arange(16, 24), np.arange(16, 24), np.arange(16, 24), indexing="ij"), -1).reshape(-1, 3)  # a full leaf
    pts = np.unique(np.concatenate(pts + [dense]), axis=0)
    vals = rng.uniform(-3, 3, len(pts)).astype(np.float32)
    return pts.astype(np.int32), vals


def expected_bits(vals, half):
    if half:
        h = vals.astype(np.float16).view(np.uint16).astype(np.uint32)
        return ((h >> 8) << 16) | ((h & 255) << 24)  # from_f16_bites, read.rs:635-642: [0, 0, b0, b1] little-endian
    return vals.astype(np.float32).view(np.uint32)


CASES = [(V.NONE, "none"), (V.ACTIVE_MASK, "mask"), (V.ZIP, "zip"), (V.ZIP | V.ACTIVE_MASK, "zip+mask"), (V.BLOSC, "blosc"),
         (V.BLOSC | V.ACTIVE_MASK, "blosc+mask")]


@pytest.mark.parametrize("comp,tag", CASES)
@pytest.mark.parametrize("half", [True, False])
@pytest.mark.parametrize("md", [0, 6])
def test_reader_recovers_written_voxels(tmp_path, comp, tag, half, md):
    pts, vals = sample_voxels()
    path = str(tmp_path / f"{tag}.vdb")
    V.VdbWriter(compression=comp, half_float=half, leaf_metadata=md).write(path, pts, vals)
    if md == 6 and comp & (V.ZIP | V.BLOSC):  # dense leaf buffers are mostly background: the codecs really ran
        import os
        plain = len(V.VdbWriter(compression=comp & V.ACTIVE_MASK, half_float=half, leaf_metadata=md).build(pts, vals))
        assert os.path.getsize(path) < plain - 500  # (node masks, which dominate this small file, are never compressed)
    r = W.VdbReader(path)
    v = r.read_vdb345_grid("ls_test")
    assert r.info.grid_compression == comp and bool(r.info.is_half_float) == half and r.info.file_voxel_count == len(pts)
    assert v.count_leaf_values() == len(pts)
    want = expected_bits(vals, half)
    for p, wbits in list(zip(pts.tolist(), want.tolist()))[::7]:
        e = v.get_voxel(p)
        assert e.kind == "Leaf" and e.value == wbits, (p, e.kind, hex(e.value), hex(wbits))
    assert v.get_voxel([16 + 8, 16, 16]).kind != "Leaf"  # just outside the dense leaf
    # the oracle's reader: same tree for everything it supports
    try:
        t, info = O.Tree.read(path, "ls_test")
    except IOError as e:  # the oracle has no Blosc decoder (the reference's is the c-blosc library): WXO_ERR_BLOSC,
        assert comp & V.BLOSC and "-7" in str(e)  # unless every block of the file happened to be stored raw
        return
    assert info.grid_compression == comp
    o = scenes.OracleScene(t, sdf=False)
    f = v.to_flat(narrow_leaves=False)
    assert np.array_equal(f.origins, o.origins)
    for a, b in ((f.kids5, o.kids5), (f.kids4, o.kids4), (f.vals3, o.vals3)):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("blocksize", [256, 384, 1024])
def test_blosc_multi_block_frames(tmp_path, blocksize):
    """Several blocks per frame, a shorter last block (never split), split and unsplit streams."""
    pts, vals = sample_voxels(seed=11)
    path = str(tmp_path / "b.vdb")
    V.VdbWriter(compression=V.BLOSC, half_float=True, leaf_metadata=6, blosc_blocksize=blocksize).write(path, pts, vals)
    v = W.VdbReader(path).read_vdb345_grid("ls_test")
    want = expected_bits(vals, True)
    for p, wbits in zip(pts.tolist(), want.tolist()):
        e = v.get_voxel(p)
        assert e.kind == "Leaf" and e.value == wbits


@pytest.mark.parametrize("codec", ["blosclz", "zlib", "lz4", "snappy", "zstd"])
@pytest.mark.parametrize("bit_shuffle", [False, True])
@pytest.mark.parametrize("blocksize,half", [(None, True), (384, True), (1024, False), (4096, False)])
def test_blosc_codecs_and_shuffles(tmp_path, codec, bit_shuffle, blocksize, half):
    """The Blosc frame with each codec c-blosc 1.x knows (BloscLZ = its default codec, LZ4 = what OpenVDB writes, Snappy, zlib, Zstd;
    the Snappy and Zstd streams are written by libsnappy / libzstd through pyarrow) and with byte or bit shuffle, single and
    multi-block, split and unsplit streams: the written voxels come back."""
    if codec in ("snappy", "zstd"):
        pytest.importorskip("pyarrow")
    pts, vals = sample_voxels(seed=5)
    path = str(tmp_path / "c.vdb")
    V.VdbWriter(compression=V.BLOSC | V.ACTIVE_MASK, half_float=half, blosc_blocksize=blocksize, blosc_codec=codec,
                blosc_bit_shuffle=bit_shuffle).write(path, pts, vals)
    v = W.VdbReader(path).read_vdb345_grid("ls_test")
    assert v.count_leaf_values() == len(pts)
    want = expected_bits(vals, half)
    for p, wbits in zip(pts.tolist(), want.tolist()):
        e = v.get_voxel(p)
        assert e.kind == "Leaf" and e.value == wbits


def test_blosclz_stream_forms():
    """The stream forms a real BloscLZ encoder emits, fed to the product's decoder through one-leaf frames: long literal runs,
    a long run (overlapping match at distance 1), length extensions of several bytes, and a far (16-bit) distance."""
    import struct
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, 9000, dtype=np.uint8).tobytes()
    cases = [bytes(700), noise[:40] + bytes([7]) * 900 + noise[40:80], noise[:300] + noise[:300],
             noise[:8500] + noise[100:500] + bytes(50), b"abc" * 400 + noise[:33], noise[:100] + bytes(70000) + noise[:100]]
    for data in cases:
        comp = V.blosclz_compress_block(data)
        assert len(comp) < len(data)
        # a one-stream Blosc frame around it (typesize 1, no shuffle, not split)
        frame = struct.pack("<BBBBIII", 2, 1, 0x10 | (0 << 5), 1, len(data), len(data), 16 + 4 + 4 + len(comp)) + struct.pack("<i", 20) + \
            struct.pack("<i", len(comp)) + comp
        out = W.vdb.blosc_decompress(frame)
        assert out == data
    far = V.blosclz_compress_block(cases[3])
    assert bytes([(7 << 5) | 31]) in far and len(far) < 8500 + 300  # the 400-byte repeat at distance 8400 went out as a far match


def test_corrupt_blosc_frame_is_an_error_not_a_crash(tmp_path):
    pts, vals = sample_voxels()
    raw = bytearray(V.VdbWriter(compression=V.BLOSC | V.ACTIVE_MASK, half_float=True).build(pts, vals))
    raw[-40] ^= 0xFF  # inside the last leaf's LZ4 stream or its header
    raw[-90] ^= 0x55
    p = tmp_path / "bad.vdb"
    p.write_bytes(bytes(raw))
    try:
        v = W.VdbReader(str(p)).read_vdb345_grid("ls_test")
        assert v.count_leaf_values() == len(pts)  # the damage hit literal bytes only: still a valid stream
    except W.vdb.VdbError as e:
        assert e.status in (-101, -107, -108) or "Blosc" in str(e)


def test_compressed_model_renders(tmp_path):
    """zlib file -> product reader -> host sweep -> flat tree == the same voxels set directly."""
    pts, vals = sample_voxels()
    path = str(tmp_path / "z.vdb")
    V.VdbWriter(compression=V.ZIP | V.ACTIVE_MASK, half_float=True).write(path, pts, vals)
    a = W.VdbReader(path).read_vdb345_grid("ls_test")
    b = W.VDB345()
    b.set_voxels(pts)
    a.compute_sdf(), b.compute_sdf()
    fa, fb = a.to_flat(False), b.to_flat(False)
    for k in ("origins", "kids5", "kids4", "vals3", "tab5", "tab4"):
        assert np.array_equal(getattr(fa, k), getattr(fb, k)), k
    inactive = ~scenes.bits2d(fa.vals3)
    assert np.array_equal(fa.tab3[inactive], fb.tab3[inactive])


@pytest.mark.parametrize("codec", ["blosclz", "lz4", "zlib", "snappy", "zstd"])
def test_mutated_blosc_frames_never_crash(codec):
    """Bit flips, truncations and random words in Blosc frames of every codec: decoded bytes or a VdbError, never a crash or an
    over-read (this test is part of the ASan + UBSan run, tools/asan_host.sh)."""
    if codec in ("snappy", "zstd"):
        pytest.importorskip("pyarrow")
    rng = np.random.default_rng(17)
    base = (rng.integers(0, 4, 6000, dtype=np.uint8).astype(np.uint16) * 257).tobytes()
    frames = [V.blosc_compress(base, 2, codec=codec), V.blosc_compress(base, 2, codec=codec, bit_shuffle=True, blocksize=1024),
              V.blosc_compress(base, 4, codec=codec, blocksize=2048)]
    ok = bad = 0
    for f in frames:
        assert W.vdb.blosc_decompress(f) == base
        for _ in range(250):
            m = bytearray(f)
            kind = rng.integers(0, 3)
            if kind == 0:
                for _ in range(int(rng.integers(1, 4))):
                    m[int(rng.integers(0, len(m)))] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1:
                m = m[:int(rng.integers(0, len(m)))]
            else:
                at = int(rng.integers(0, max(1, len(m) - 4)))
                m[at:at + 4] = rng.integers(0, 256, 4, dtype=np.uint8).tobytes()
            try:
                W.vdb.blosc_decompress(bytes(m))
                ok += 1
            except W.vdb.VdbError:
                bad += 1
    assert ok + bad == 750 and bad > 100


def test_crafted_blosc_header_is_rejected_before_any_allocation():
    """A 16-byte frame whose header claims 4 GiB: the decoder must refuse from the header alone -- the size is
    checked against what the caller expects / has room for, and against what the stored bytes could possibly expand to."""
    import ctypes as C
    import time
    from woxel_b200 import _ffi
    lib = _ffi.host_lib()
    for flags in (0x21, 0x01, 0x61, 0x25):  # LZ4 / BloscLZ / zlib with byte shuffle, LZ4 with bit shuffle
        frame = bytes([2, 1, flags, 4]) + (0xFFFFFFF0).to_bytes(4, "little") + (0xFFFFFFF0).to_bytes(4, "little") + (16).to_bytes(4, "little")
        out = C.create_string_buffer(1 << 16)
        n = C.c_size_t(0)
        t0 = time.perf_counter()
        rc = lib.wxh_blosc_decompress(frame, len(frame), out, len(out), C.byref(n))
        assert rc != 0 and time.perf_counter() - t0 < 0.05, (flags, rc)
        # the same header with room for it: still refused (16 stored bytes cannot expand to 4 GiB)
        frame2 = frame + b"\0" * 64
        frame2 = frame2[:12] + len(frame2).to_bytes(4, "little") + frame2[16:]
        t0 = time.perf_counter()
        with pytest.raises(W.vdb.VdbError):
            W.vdb.blosc_decompress(frame2[:4] + (1 << 20).to_bytes(4, "little") + (0xFFFFFFF0).to_bytes(4, "little") + frame2[12:])
        assert time.perf_counter() - t0 < 0.5


def test_blosc_frames_with_streams_from_the_real_codec_libraries():
    """tests/golden/blosc_frames.npz (tests/golden/make_blosc_golden.py): frames whose compressed streams were produced by liblz4,
    libsnappy, libzstd (all three through pyarrow) and zlib -- not by this repository's own encoders -- in every shuffle / block-size
    combination.  The product decoder must return the payloads byte for byte.  (The frame CONTAINER is still written by
    tests/vdb_writer.py: c-blosc itself is not available in this image.)"""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "blosc_frames.npz"))
    frames = [k for k in z.files if k.startswith("frame/")]
    assert len(frames) >= 130
    seen = set()
    for k in frames:
        _, pname, codec, shuffle, _ = k.split("/")
        assert W.vdb.blosc_decompress(z[k].tobytes()) == z[f"payload/{pname}"].tobytes(), k
        seen.add((codec, shuffle))
    assert seen >= {(c, s) for c in ("lz4", "snappy", "zlib", "zstd1", "zstd19") for s in ("none", "byte", "bit")}
    # bare Zstandard frames of libzstd (levels -5 .. 22; multi-block inputs, RLE / raw / compressed blocks, 1- and 4-stream
    # Huffman literals, FSE-coded and repeated tables) inside a one-stream Blosc frame
    raw = [k for k in z.files if k.startswith("raw_zstd/")]
    assert len(raw) >= 40
    for k in raw:
        data = z["payload/" + k.split("/")[1]].tobytes()
        stream = z[k].tobytes()
        frame = V.blosc_compress(data, 1, do_shuffle=False, codec="zstd", encode=lambda b: stream)
        assert W.vdb.blosc_decompress(frame) == data, k
    # a real-library stream cut short or with a flipped byte is an error, never a crash or a wrong size
    k = "frame/sdf_f32/lz4/byte/0"
    good = bytearray(z[k].tobytes())
    for cut in (len(good) - 1, len(good) // 2, 40):
        with pytest.raises(W.vdb.VdbError):
            W.vdb.blosc_decompress(bytes(good[:cut]))
