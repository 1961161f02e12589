#!/bin/bash
# BASELINE.json configs[3]: the reference's own 3-D Harris reconnection deck (sample/reconnection/reconnection) on N
# ranks, one GPU per rank, hot path on the GPUs through LD_PRELOAD (the drop-in multi-rank seam: device boundary_p and
# field kernels, exchange through the host program's mp_* ports over oracle/mpi_shim's shared-memory MPI).
#   tools/c4_run.sh N NX NY NZ NPPC STEPS [cpu]      "cpu" = the same run without the preload (the CPU reference)
N=$1; NX=$2; NY=$3; NZ=$4; NPPC=$5; STEPS=$6; MODE=${7:-gpu}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
D=$(mktemp -d /tmp/c4.XXXXXX)
cd "$D" || exit 1
export VPIC_REC_NX=$NX VPIC_REC_NY=$NY VPIC_REC_NZ=$NZ VPIC_REC_NPPC=$NPPC VPIC_REC_STEPS=$STEPS VPIC_REC_RESTART=100000000
export VPIC_SHIM_HEAP_MB=${VPIC_SHIM_HEAP_MB:-256}
if [ "$MODE" = gpu ]; then
  export LD_PRELOAD=$ROOT/vpic_b200/libvpic_b200.so VPIC_B200_TRACE=1 VPIC_B200_STRICT=1
  GFLAG=-g; [ -n "$C4_SHARE_GPU" ] && GFLAG=; TPP=1
else
  GFLAG=; TPP=$(( $(nproc) / N )); [ $TPP -lt 1 ] && TPP=1
fi
t0=$(date +%s.%N)
$ROOT/oracle/mpi_shim/shimrun -n $N $GFLAG $ROOT/oracle/_ref/reconnection.scalar --tpp $TPP > out.txt 2>&1
rc=$?
t1=$(date +%s.%N)
unset LD_PRELOAD
python3 - "$N" "$NX" "$NY" "$NZ" "$NPPC" "$STEPS" "$MODE" "$rc" "$t0" "$t1" <<'PY'
import sys, re, json
n, nx, ny, nz, nppc, steps, mode, rc, t0, t1 = sys.argv[1:]
out = open("out.txt").read()
prof = {}
for ln in out.splitlines():
    m = re.match(r"\s*(\w+) \|\s*\d+% ([0-9.e+-]+) ([0-9.e+-]+) ([0-9.e+-]+) \|", ln)
    if m:
        prof[m.group(1)] = (float(m.group(2)), float(m.group(3)))
npart = 2.0 * float(nx) * float(ny) * float(nz) * float(nppc)
res = {"deck": "sample/reconnection/reconnection", "ranks": int(n), "mode": mode, "rc": int(rc), "grid": [int(nx), int(ny), int(nz)],
       "nppc": float(nppc), "steps": int(steps), "particles": npart, "wall_s": float(t1) - float(t0),
       "normal_exit": "normal exit" in out}
# the reference's own profile table, rank 0: seconds spent in each stage of the step loop
keep = ("advance_p", "boundary_p", "sort_p", "clear_accumulators", "reduce_accumulators", "unload_accumulator", "advance_b",
        "advance_e", "load_interpolator", "synchronize_jf", "clean_div_e", "clean_div_b", "user_diagnostics", "user_particle_injection")
res["profile_s"] = {k: prof[k][0] for k in keep if k in prof}
loop = sum(v for k, v in res["profile_s"].items() if k not in ("user_diagnostics",))
if loop > 0:
    res["step_loop_s"] = loop
    res["pushes_per_s"] = npart * int(steps) / loop
tr = [ln for ln in out.splitlines() if ln.startswith("vpic_b200 trace[0]")]
if tr:
    res["trace_rank0"] = tr[-1][:600]
    hs = [ln for ln in out.splitlines() if ln.startswith("vpic_b200 host seconds[0]")]
    if hs:
        res["host_seconds_rank0"] = hs[-1]
try:
    en = [ln.split() for ln in open("rundata/energies") if not ln.startswith("%")]
    res["energies_first_last"] = [en[0], en[-1]]
except Exception:
    pass
print(json.dumps(res))
PY
tail -3 out.txt | cut -c1-200 >&2
cd /; rm -rf "$D"
