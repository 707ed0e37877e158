set -e
for cfg in "8 2" "12 2"; do
  set -- $cfg
  LDPC_B200_DEFINES="-DLDPC_I8_WARPS=$1 -DLDPC_I8_MINBLOCKS=$2" python ldpc_toolbox_b200/build.py --force > /dev/null
  echo "== warps=$1 minblocks=$2"
  python tools/quick_bench.py --tiles 1184 --iters 10 --mean 2.24 --std 2.12 --signs 1 --reps 2 | cut -c1-200
done
