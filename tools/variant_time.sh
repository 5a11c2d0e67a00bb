#!/bin/bash
# Times the 1M-floe contact step with the product library and with every build_exp/<name> given (tools/variant_build.sh),
# through the stand-alone C++ driver (no Python, ~4 s per run): step, phases, class C kernel, force sums (a coarse parity check)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name, libdir
  out=$(LD_LIBRARY_PATH=$2 timeout 120 tools/sz_driver 1000000 ${STEPS:-6} 3 2>&1 | tail -1)
  python3 - "$1" "$out" <<'P'
import sys, json
name, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    print("%-12s step %.3f  broad %.3f narrow %.3f assembly %.3f  C %.3f  sum_fx %.6e sum_fy %.6e" % (name, d["ms_per_step"], d["phase_ms"]["broad"], d["phase_ms"]["narrow"], d["phase_ms"]["assembly"], d["narrow_class_C"]["ms"], d["sum_fx"], d["sum_fy"]))
except Exception as e:
    print(name, "FAILED", line[-300:])
P
}
{
run default subzero_b200/_lib
for v in "$@"; do run $v build_exp/$v; done
run default2 subzero_b200/_lib
} 2>&1 | tee gpurun_out/variant_time_${TAG:-x}.log
