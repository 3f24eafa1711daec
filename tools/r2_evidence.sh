#!/bin/sh
# Round-2 evidence on one B200 (run under gpurun): GPU tests, smoke, both bench arms, ncu launch list of the bench,
# ncu --set full of every kernel besides the scan kernel (that one: tools/ncu_target.py 4 1, profiles/r02_scan_cells_*),
# one compute-sanitizer memcheck pass.  Outputs land in gpurun_out/ and are summarised into profiles/ afterwards.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1; tail -3 $O/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; tail -1 $O/r02_smoke.log
python bench.py --impl reference > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err
python bench.py > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err; cut -c1-200 $O/r02_bench_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cli --pruned-perms 60 --e2e-perms 2 > $O/r02_launches_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_cells -s 3 -c 1 -o $O/r02_scan_c4 -f \
    python tools/ncu_target.py 4 1 > $O/ncu_r02_scan_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bucket_count -s 1 -c 1 -o $O/r02_bucket_count -f \
    python tools/ncu_target.py 4 1 > $O/ncu_r02_bucket_count.log 2>&1
if [ "$1" = "all" ]; then      # the captures of the other kernels are 57 MB: on request only (gpurun returns at most 64 MB)
python tools/ncu_target.py 2 4 300 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on \
    -k regex:'gen_blocks|permute_kernel|bucket|taskgen|replay|compact|adjust_background|pjh|pbk|knots|spline' -c 40 \
    -o $O/r02_other_kernels -f python tools/ncu_target.py 2 4 300 > $O/r02_ncu_other.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'pack_clr|project' -c 2 \
    -o $O/r02_tail_kernels -f python tools/ncu_target.py 2 4 300 > $O/r02_ncu_tail.log 2>&1
fi
# (compute-sanitizer is closed on this pool: profiles/r02_memcheck.txt holds its answer)
