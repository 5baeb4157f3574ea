"""Generates tests/golden/blosc_frames.npz: c-blosc-1 frames whose compressed STREAMS come from the real codec libraries --
liblz4 (pyarrow codec "lz4_raw": the LZ4 block format c-blosc's LZ4 codec writes and OpenVDB uses), libsnappy (pyarrow "snappy":
the raw Snappy format of c-blosc's codec 2), libzstd (pyarrow "zstd": the Zstandard frames of c-blosc's codec 4) and zlib -- wrapped in the frame container by tests/vdb_writer.blosc_compress.
c-blosc itself (blosc-src 0.2.1 in the reference's Cargo.lock, `read.rs:514-533`) is not available here, so the container is still
written by this repository; the codec streams are not.  Run in the build container (pyarrow is in the image):

    python tests/golden/make_blosc_golden.py
"""
import os
import sys
import zlib

import numpy as np
import pyarrow as pa

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import vdb_writer as VW  # noqa: E402

rng = np.random.default_rng(20261017)
payloads = {
    "sdf_f32": (np.clip(np.cumsum(rng.normal(0, 0.02, 4096)), -3, 3).astype(np.float32).tobytes(), 4),      # smooth level-set values
    "half_u16": (np.repeat(rng.integers(0, 2000, 300), 7)[:2048].astype(np.uint16).tobytes(), 2),          # runs, half floats
    "mask_bytes": (np.packbits(rng.random(8 * 4096) < 0.1).tobytes() + b"\0" * 777, 1),                       # sparse masks, odd tail
    "noise_f32": (rng.random(1500).astype(np.float32).tobytes(), 4),                                         # barely compressible
    "ramp_i64": (np.arange(-700, 1348, dtype=np.int64).tobytes(), 8),
}
# larger inputs for the bare Zstandard streams (several 128 KiB blocks: repeated tables, tree reuse, RLE and raw blocks)
zstd_payloads = {
    "nibbles_140k": bytes(rng.integers(0, 16, 140000, dtype=np.uint8)),
    "digits_text": b"".join(bytes(str(i * i % 977), "ascii") + b"," for i in range(60000)),
    "zeros_300k": b"\0" * 300000,
    "mixed": b"".join((bytes(rng.integers(0, 256, int(rng.integers(1, 3000)), dtype=np.uint8)) if rng.random() < 0.3 else
                       bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 5000)) if rng.random() < 0.5 else
                       (b"voxel%d" % int(rng.integers(0, 50))) * int(rng.integers(1, 300))) for _ in range(120)),
}
enc = {
    "lz4": lambda b: pa.compress(b, codec="lz4_raw", asbytes=True),
    "snappy": lambda b: pa.compress(b, codec="snappy", asbytes=True),
    "zlib": lambda b: zlib.compress(b, 9),
    "zstd1": lambda b: pa.Codec("zstd", compression_level=1).compress(b, asbytes=True),
    "zstd19": lambda b: pa.Codec("zstd", compression_level=19).compress(b, asbytes=True),
}
out = {}
for pname, (data, typesize) in payloads.items():
    out[f"payload/{pname}"] = np.frombuffer(data, np.uint8)
    for codec, fn in enc.items():
        for shuffle in ("none", "byte", "bit"):
            for blocksize in (None, 2048):
                if shuffle == "bit" and typesize == 1:
                    continue
                frame = VW.blosc_compress(data, typesize, do_shuffle=(shuffle == "byte"), blocksize=blocksize, codec=("zstd" if codec.startswith("zstd") else codec),
                                          bit_shuffle=(shuffle == "bit"), encode=fn)
                out[f"frame/{pname}/{codec}/{shuffle}/{blocksize or 0}"] = np.frombuffer(frame, np.uint8)
# bare streams too (no container): what the codec libraries themselves produced
for pname, (data, _) in payloads.items():
    out[f"raw_lz4/{pname}"] = np.frombuffer(enc["lz4"](data), np.uint8)
    out[f"raw_snappy/{pname}"] = np.frombuffer(enc["snappy"](data), np.uint8)
for pname, data in list(zstd_payloads.items()) + [(k, v[0]) for k, v in payloads.items()]:
    out[f"payload/{pname}"] = np.frombuffer(data, np.uint8)
    for level in ((-5, 1, 3, 7, 12, 19, 22) if len(data) < 100000 else (1, 5, 19)):
        out[f"raw_zstd/{pname}/{level}"] = np.frombuffer(pa.Codec("zstd", compression_level=level).compress(data, asbytes=True), np.uint8)
np.savez_compressed(os.path.join(HERE, "blosc_frames.npz"), **out)
print(len(out), "arrays,", sum(v.nbytes for v in out.values()), "bytes; pyarrow", pa.__version__)
