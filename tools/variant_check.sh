# Pre-flight of a build-time variant of the raycast kernel, no GPU needed:
#   bash tools/variant_check.sh NAME "-DWX_SOMETHING [-D...]"
# 1. builds woxel_b200/libwoxel_b200_NAME.so with the flags (sm_100a);  2. prints registers / spills / SASS size of
# raycast_kernel<0,false> next to the default build's;  3. runs the CPU emulation tests (tests/test_device_emu.py) with the
# same flags: the variant must stay bit-identical to the oracle.  Then time it:  gpurun -- "bash tools/gpu/ab.sh default NAME"
set -e
cd "$(dirname "$0")/.."
NAME=$1
FLAGS=$2
[ -n "$NAME" ] || { echo "usage: $0 NAME \"-D...\""; exit 2; }
K='_ZN2wx14raycast_kernelILi0ELb0EEEvNS_12RenderParamsE'
sass_lines() { cuobjdump -sass -fun "$K" "$1" 2>/dev/null | grep -cE '^\s+/\*[0-9a-f]{4}\*/'; }
make -C woxel_b200/csrc > /dev/null
cp woxel_b200/csrc/build.log /tmp/build_default.log
make -C woxel_b200/csrc EXTRA="$FLAGS" OUT=../libwoxel_b200_$NAME.so > /dev/null
cp woxel_b200/csrc/build.log /tmp/build_$NAME.log
for v in default $NAME; do
  so=woxel_b200/libwoxel_b200.so; [ $v = default ] || so=woxel_b200/libwoxel_b200_$NAME.so
  echo "$v: $(sass_lines $so) SASS instructions; $(grep -A3 "Compiling entry function '$K'" /tmp/build_$v.log | grep -oE 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ',' )"
done
make -C woxel_b200/csrc > /dev/null   # leave build.log describing the default build
WX_EMU_EXTRA="$FLAGS" python -m pytest tests/test_device_emu.py -q -x -p no:cacheprovider 2>&1 | tail -2
