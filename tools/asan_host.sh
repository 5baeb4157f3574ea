# AddressSanitizer + UBSan over the C++ host library (the .vdb reader with its zlib / Blosc / LZ4 decoders, the VDB345 tree,
# compute_sdf, to_flat, ComputeState::build) and over the oracle's C restatement, driven by the CPU tests.  No GPU needed.
#   bash tools/asan_host.sh
set -e
cd "$(dirname "$0")/.."
make -C woxel_b200/host asan > /dev/null
make -C oracle libwxo_asan.so > /dev/null
export WOXEL_HOST_LIB=$PWD/woxel_b200/libwoxel_host_asan.so
export WXO_VARIANT=asan
export LD_PRELOAD="$(/usr/bin/gcc -print-file-name=libasan.so) $(/usr/bin/gcc -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=1:handle_segv=0
export UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
python -m pytest -s tests/test_reader_fuzz.py tests/test_vdb_compressed.py tests/test_host_vs_oracle.py tests/test_oracle_reference_vectors.py tests/test_scenegen.py -q -m "not gpu" -p no:cacheprovider "$@"
# Second stage: the DEVICE code of woxel_b200/csrc/wx_device.cuh + the tree packing of wx_pack.h, compiled for the host (tests/emu)
# with the same sanitizers and driven by the emulation tests (table walks, both marches, shading, the work-queue protocol).
export CXX=/usr/bin/g++
export WX_EMU_EXTRA="-fsanitize=address,undefined -fno-sanitize-recover=undefined -g"
python -m pytest -s tests/test_device_emu.py -q -m "not gpu" -p no:cacheprovider -k "not full_size" "$@"
