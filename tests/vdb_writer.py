"""Test-side writer of OpenVDB files (TEST INFRASTRUCTURE, not product code, not a restatement of the reference's
writer -- src/vdb/write.rs is out of scope and has known endianness bugs).

It emits what the reference's READER expects (src/vdb/read.rs:62-629): file version 224, one
`Tree_float_5_4_3[_HalfFloat]` grid, per-grid compression flags NONE / ACTIVE_MASK / ZIP / BLOSC in any combination
the format allows.  Blosc blocks are c-blosc 1.x frames (16-byte header, block offsets, per-block streams split by
type size, byte shuffle) holding LZ4 block streams -- the container OpenVDB writes with
blosc_compress_ctx(9, shuffle, sizeof(T), ..., "lz4") -- produced by the small encoder below, written from the
published formats (c-blosc README_HEADER / blosc.c, lz4 Block Format); no c-blosc binary exists in this image to
cross-check against.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

NONE, ZIP, ACTIVE_MASK, BLOSC = 0, 1, 2, 4


# ------------------------------------------------------------------------------------------------
# LZ4 block format + c-blosc 1.x frame
# ------------------------------------------------------------------------------------------------
def lz4_compress_block(src: bytes) -> bytes:
    """Greedy LZ4 block encoder (4-byte hash matches, 64 KB window).  Valid per the LZ4 block format: the last 5 bytes
    are literals and the last match starts at least 12 bytes before the end."""
    n = len(src)
    out = bytearray()
    table: dict[bytes, int] = {}
    anchor = i = 0
    limit = n - 12

    def emit(lit_end: int, match_len: int, offset: int):
        nonlocal anchor
        lit = src[anchor:lit_end]
        ll, ml = len(lit), match_len - 4 if match_len else 0
        token = (min(ll, 15) << 4) | (min(ml, 15) if match_len else 0)
        out.append(token)
        if ll >= 15:
            r = ll - 15
            while r >= 255:
                out.append(255)
                r -= 255
            out.append(r)
        out.extend(lit)
        if match_len:
            out.extend(struct.pack("<H", offset))
            if ml >= 15:
                r = ml - 15
                while r >= 255:
                    out.append(255)
                    r -= 255
                out.append(r)

    while i < limit:
        key = src[i:i + 4]
        cand = table.get(key)
        table[key] = i
        if cand is not None and i - cand <= 65535:
            m = 4
            while i + m < n - 5 and src[cand + m] == src[i + m]:
                m += 1
            emit(i, m, i - cand)
            i += m
            anchor = i
        else:
            i += 1
    emit(n, 0, 0)
    return bytes(out)


def blosclz_compress_block(src: bytes) -> bytes:
    """Greedy encoder of the BloscLZ / FastLZ level-2 stream format (3-byte hash matches): literal runs of 1..32 bytes
    (control byte = count - 1), matches of 3.. bytes (control byte = min(len - 2, 7) << 5 | distance high bits, length
    extension bytes, distance low byte; distances of 8192.. as the far form: 31 / 255 / 16-bit big-endian distance - 8192).
    The stream starts with a literal run, as the format requires."""
    n = len(src)
    out = bytearray()
    table: dict[bytes, int] = {}
    lits = bytearray()

    def flush_literals():
        nonlocal lits
        for k in range(0, len(lits), 32):
            run = lits[k:k + 32]
            out.append(len(run) - 1)
            out.extend(run)
        lits = bytearray()

    i = 0
    while i < n:
        cand = table.get(src[i:i + 3]) if i + 3 <= n else None
        if i + 3 <= n:
            table[src[i:i + 3]] = i
        dist = i - cand if cand is not None else 0
        if cand is not None and i > 0 and dist <= 65535 + 8191 and (len(lits) > 0 or len(out) > 0):
            m = 3
            while i + m < n and src[cand + m] == src[i + m]:
                m += 1
            flush_literals()
            v = m - 2
            d = dist - 1
            far = d >= 8191
            hi = 31 if far else d >> 8
            if v < 7:
                out.append((v << 5) | hi)
            else:
                out.append((7 << 5) | hi)
                r = v - 7
                while r >= 255:
                    out.append(255)
                    r -= 255
                out.append(r)
            if far:
                out.append(255)
                out.extend(struct.pack(">H", d - 8191))
            else:
                out.append(d & 255)
            i += m
        else:
            lits.append(src[i])
            i += 1
    flush_literals()
    return bytes(out)


def bitshuffle(data: bytes, typesize: int) -> bytes:
    """c-blosc's bit shuffle of one block: the first (n // 8) * 8 elements become typesize * 8 bit rows (row j * 8 + b = bit b of
    byte j of every element, element 8k + m in bit m of byte k of the row); the rest of the block is left as it is."""
    a = np.frombuffer(data, np.uint8)
    ne8 = (len(a) // typesize) // 8 * 8
    head = a[:ne8 * typesize].reshape(ne8, typesize)
    bits = np.unpackbits(head[:, :, None], axis=2, bitorder="little")          # [element, byte j, bit b]
    rows = bits.transpose(1, 2, 0).reshape(typesize * 8, ne8)                  # row j * 8 + b, one bit per element
    packed = np.packbits(rows, axis=1, bitorder="little")                      # element 8k + m -> bit m of byte k
    return packed.tobytes() + a[ne8 * typesize:].tobytes()


def shuffle(data: bytes, typesize: int) -> bytes:
    a = np.frombuffer(data, np.uint8)
    n = len(a) // typesize
    head = a[:n * typesize].reshape(n, typesize).T.reshape(-1)
    return head.tobytes() + a[n * typesize:].tobytes()


BLOSC_CODECS = {"blosclz": 0, "lz4": 1, "snappy": 2, "zlib": 3, "zstd": 4}


def _pyarrow_codec(name: str):
    """Snappy and Zstd have no encoder in this module: their streams come from libsnappy / libzstd through pyarrow."""
    import pyarrow as pa
    return lambda b: pa.compress(b, codec=name, asbytes=True)


def blosc_compress(data: bytes, typesize: int, do_shuffle: bool = True, blocksize: int | None = None, force_memcpy: bool = False,
                   codec: str = "lz4", bit_shuffle: bool = False, encode=None) -> bytes:
    """A c-blosc 1.x frame with the LZ4 codec.  header: version 2, versionlz 1, flags (bit0 shuffle, bit1 memcpyed,
    bits 5-7 = 1: LZ4), typesize, nbytes, blocksize, cbytes; then int32 block offsets; a block is split into `typesize`
    streams when typesize <= 16 and blocksize / typesize >= 128 (never the leftover block); every stream is an int32
    length followed by that many bytes (length == stream size: stored raw)."""
    nbytes = len(data)
    if blocksize is None:
        blocksize = max(nbytes, 1)
    split = typesize <= 16 and blocksize // typesize >= 128
    byte_sh = do_shuffle and typesize > 1 and not bit_shuffle
    flags = (1 if byte_sh else 0) | (4 if bit_shuffle else 0) | (0 if split else 0x10) | (BLOSC_CODECS[codec] << 5)
    # `encode`: the stream compressor; by default this module's own encoders -- tests/golden/make_blosc_golden.py passes the REAL
    # libraries' (liblz4 / libsnappy through pyarrow, zlib) so that the decoders are also checked against streams they did not write
    if encode is None:
        encode = {"lz4": lz4_compress_block, "blosclz": blosclz_compress_block, "zlib": lambda b: zlib.compress(b, 6),
                  "snappy": lambda b: _pyarrow_codec("snappy")(b), "zstd": lambda b: _pyarrow_codec("zstd")(b)}[codec]
    if force_memcpy or nbytes < 128:
        flags |= 2
        return struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, blocksize, 16 + nbytes) + data
    nblocks = (nbytes + blocksize - 1) // blocksize
    body = bytearray()
    starts = []
    for b in range(nblocks):
        starts.append(16 + 4 * nblocks + len(body))
        chunk = data[b * blocksize:(b + 1) * blocksize]
        leftover = len(chunk) < blocksize
        if flags & 1:
            chunk = shuffle(chunk, typesize)
        elif flags & 4 and blocksize >= typesize:
            chunk = bitshuffle(chunk, typesize)
        nsplits = typesize if (split and not leftover) else 1
        neblock = len(chunk) // nsplits
        for k in range(nsplits):
            part = chunk[k * neblock:(k + 1) * neblock] if nsplits > 1 else chunk
            comp = encode(part)
            if len(comp) >= len(part):
                comp = part  # stored: a stream as long as its block is a plain copy
            body += struct.pack("<i", len(comp)) + comp
    frame = bytearray(struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, blocksize, 0))
    for s in starts:
        frame += struct.pack("<i", s)
    frame += body
    struct.pack_into("<I", frame, 12, len(frame))
    return bytes(frame)


# ------------------------------------------------------------------------------------------------
# .vdb
# ------------------------------------------------------------------------------------------------
def _len_str(s: str) -> bytes:
    b = s.encode()
    return struct.pack("<I", len(b)) + b


def _meta(entries) -> bytes:
    out = struct.pack("<I", len(entries))
    for name, typ, payload in entries:
        out += _len_str(name) + _len_str(typ) + struct.pack("<I", len(payload)) + payload
    return out


def _mask_bytes(bits: np.ndarray) -> bytes:
    return np.packbits(bits.astype(np.uint8), bitorder="little").tobytes()


class VdbWriter:
    def __init__(self, compression: int = ACTIVE_MASK, half_float: bool = True, leaf_metadata: int = 0, blosc_blocksize: int | None = None,
                 blosc_codec: str = "lz4", blosc_bit_shuffle: bool = False):
        self.compression, self.half, self.md, self.blosc_blocksize = compression, half_float, leaf_metadata, blosc_blocksize
        self.blosc_codec, self.blosc_bit_shuffle = blosc_codec, blosc_bit_shuffle

    # read.rs:490-574 mirrored
    def _blocks(self, raw: bytes, elem: int) -> bytes:
        if self.compression & BLOSC:
            if len(raw) == 0:
                return struct.pack("<q", 0)
            frame = blosc_compress(raw, elem, blocksize=self.blosc_blocksize, codec=self.blosc_codec, bit_shuffle=self.blosc_bit_shuffle)
            if len(frame) >= len(raw) + 16 and not self.blosc_blocksize:
                return struct.pack("<q", -len(raw)) + raw  # OpenVDB stores incompressible data raw with a negative size
            return struct.pack("<q", len(frame)) + frame
        if self.compression & ZIP:
            z = zlib.compress(raw, 6)
            if len(raw) == 0 or len(z) >= len(raw):
                return struct.pack("<q", -len(raw)) + raw
            return struct.pack("<q", len(z)) + z
        return raw

    def _values(self, vals: np.ndarray, mask: np.ndarray, md: int) -> bytes:
        """One read_compressed record (read.rs:378-488): metadata byte, [inactive values], [selection mask], blocks."""
        out = struct.pack("<B", md)
        if md in (2, 4):
            out += struct.pack("<I", 0)
        elif md == 5:
            out += struct.pack("<II", 0, 0)
        if md in (3, 4, 5):
            out += bytes(len(mask) // 8)
        keep = vals[mask] if (self.compression & ACTIVE_MASK) and md != 6 else vals
        raw = keep.astype(np.float16).tobytes() if self.half else keep.astype(np.float32).tobytes()
        return out + self._blocks(raw, 2 if self.half else 4)

    def build(self, voxels: np.ndarray, values: np.ndarray | None = None, grid_name: str = "ls_test") -> bytes:
        voxels = np.asarray(voxels, np.int64).reshape(-1, 3)
        if values is None:
            values = np.linspace(-1.0, 1.0, len(voxels)).astype(np.float32)
        # tree: N5 origin -> N4 offset -> leaf offset -> {voxel offset: value}
        tree: dict = {}
        for (x, y, z), v in zip(voxels.tolist(), values.tolist()):
            o5 = ((x >> 12) << 12, (y >> 12) << 12, (z >> 12) << 12)
            k5 = (((x & 4095) >> 7) << 10) | (((y & 4095) >> 7) << 5) | ((z & 4095) >> 7)
            k4 = (((x & 127) >> 3) << 8) | (((y & 127) >> 3) << 4) | ((z & 127) >> 3)
            k3 = ((x & 7) << 6) | ((y & 7) << 3) | (z & 7)
            tree.setdefault(o5, {}).setdefault(k5, {}).setdefault(k4, {})[k3] = v
        topo, leaves = bytearray(), bytearray()
        topo += struct.pack("<I", 1) + struct.pack("<f", 3.0) + struct.pack("<II", 0, len(tree))  # buffers, background, tiles, nodes
        n_active = 0
        for o5 in sorted(tree):
            n5 = tree[o5]
            topo += struct.pack("<iii", *o5)
            kid = np.zeros(32768, bool)
            kid[list(n5)] = True
            topo += _mask_bytes(kid) + _mask_bytes(np.zeros(32768, bool))
            topo += self._values(np.zeros(32768, np.float32), np.zeros(32768, bool), 0)
            for k5 in sorted(n5):
                n4 = n5[k5]
                kid4 = np.zeros(4096, bool)
                kid4[list(n4)] = True
                topo += _mask_bytes(kid4) + _mask_bytes(np.zeros(4096, bool))
                topo += self._values(np.zeros(4096, np.float32), np.zeros(4096, bool), 0)
                for k4 in sorted(n4):
                    leaf = n4[k4]
                    m = np.zeros(512, bool)
                    m[list(leaf)] = True
                    vals = np.full(512, 3.0, np.float32)
                    for k3, v in leaf.items():
                        vals[k3] = v
                    topo += _mask_bytes(m)
                    leaves += _mask_bytes(m) + self._values(vals, m, self.md)
                    n_active += int(m.sum())
        grid_type = "Tree_float_5_4_3" + ("_HalfFloat" if self.half else "")
        gmeta = _meta([("class", "string", b"level set"), ("file_voxel_count", "int64", struct.pack("<q", n_active)),
                       ("is_saved_as_half_float", "bool", b"\x01" if self.half else b"\x00"), ("name", "string", grid_name.encode())])
        transform = _len_str("UniformScaleMap") + struct.pack("<15d", *([1.0] * 15))
        grid_body = struct.pack("<I", self.compression) + gmeta + transform
        head = struct.pack("<Q", 0x56444220) + struct.pack("<III", 224, 10, 0) + b"\x01" + b"0" * 36 + _meta([]) + struct.pack("<I", 1)
        desc = _len_str(grid_name) + _len_str(grid_type) + _len_str("")
        grid_pos = len(head) + len(desc) + 24
        block_pos = grid_pos + len(grid_body) + len(topo)
        end_pos = block_pos + len(leaves)
        return head + desc + struct.pack("<QQQ", grid_pos, block_pos, end_pos) + grid_body + bytes(topo) + bytes(leaves)

    def write(self, path: str, voxels, values=None, grid_name: str = "ls_test") -> None:
        with open(path, "wb") as f:
            f.write(self.build(voxels, values, grid_name))
