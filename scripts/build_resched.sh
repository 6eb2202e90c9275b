#!/bin/bash
# Build the library with the post-pass SASS scheduler (tools/sass_resched.py) applied to the named kernels:
#   bash scripts/build_resched.sh _rs [extra nvcc flags ...] -- 'KERNEL_REGEX' ['KERNEL_REGEX' ...]
# -> agilex-ntt_b200/lib/libagxntt_rs.so.  nvcc's own build steps are taken from `nvcc -dryrun --keep` and run one by one;
# between ptxas and fatbinary the cubin is re-scheduled in place.  Everything else is the recipe of build.py.
set -e
suffix=$1; shift
flags=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do flags+=("$1"); shift; done
[ "$1" == "--" ] && shift
kernels=()
for k in "$@"; do kernels+=(--kernel "$k"); done
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
KEEP=$(mktemp -d /tmp/agx_resched.XXXXXX)
cd "$ROOT/agilex-ntt_b200/csrc"
mkdir -p ../lib
OUT="$ROOT/agilex-ntt_b200/lib/libagxntt$suffix.so"
nvcc -dryrun --keep --keep-dir "$KEEP" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
     "${flags[@]}" -shared -o "$OUT" agx_api.cu agx_tables.cpp 2>&1 | sed 's/^#\$ //' > "$KEEP/steps.sh"
{
  echo "set -e"
  while IFS= read -r line; do
    case "$line" in
      rm\ *) ;;                                   # keep the intermediates until the end
      ptxas\ *)
        echo "$line"
        echo "python \"$ROOT/tools/sass_resched.py\" \"$KEEP/agx_api.cubin\" \"$KEEP/agx_api.rs.cubin\" ${kernels[*]@Q} ${RESCHED_ARGS:-}"
        echo "cp \"$KEEP/agx_api.cubin\" \"$KEEP/agx_api.orig.cubin\"; cp \"$KEEP/agx_api.rs.cubin\" \"$KEEP/agx_api.cubin\""
        ;;
      *)
        if [[ "$line" =~ ^[A-Za-z_][A-Za-z0-9_]*= ]]; then   # nvcc's environment lines: keep the single-word ones (PATH, CICC_PATH ...)
          if [[ "$line" =~ ^[A-Za-z_][A-Za-z0-9_]*=[^[:space:]]*[[:space:]]*$ ]]; then echo "export $line"; fi
        else
          echo "$line"
        fi ;;
    esac
  done < "$KEEP/steps.sh"
} > "$KEEP/run.sh"
bash "$KEEP/run.sh"
echo "built $OUT  (intermediates in $KEEP)"
