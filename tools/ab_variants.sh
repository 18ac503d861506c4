#!/bin/bash
# A/B harness: builds libphx variants with different -D knobs (CPU box), then -- on the GPU
# box -- times bench.py for each.  Usage:
#   tools/ab_variants.sh build  name1="-DSC_X=1" name2="-DSC_Y=2" ...
#   tools/ab_variants.sh run [steps]        (on the GPU box; prints one line per variant)
set -e
cd "$(dirname "$0")/.."
mode=$1; shift
if [ "$mode" = build ]; then
  mkdir -p build/variants
  rm -f build/variants/*.so
  for kv in "$@"; do
    name="${kv%%=*}"; flags="${kv#*=}"
    make -s -C phantom_b200/csrc OBJDIR=../../build/variants/obj_$name OUT=../../build/variants/libphx_$name.so EXTRA="$flags" >/dev/null
    echo "built $name ($flags)"
  done
else
  steps=${1:-500}
  for so in build/variants/*.so; do
    PHX_LIB=$PWD/$so python bench.py --steps $steps --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$(basename $so)', 'us/launch=%.2f' % (d['ms_per_step']*1e3), 'frac=%.3f' % d['roofline']['frac'], 'sm_mhz', d['clocks']['sm_mhz'])"
  done
fi
