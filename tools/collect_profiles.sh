#!/bin/bash
# Runs ON THE GPU BOX (gpurun -- 'bash tools/collect_profiles.sh TAG'): bench lines, ncu launch list, one ncu --set full
# capture per GEMM kernel.  Everything lands in gpurun_out/; tools/summarize_profiles.py (run in the build container)
# turns it into the tracked summaries under profiles/.  Each ncu command runs only after the same command line has
# exited 0 without ncu (B200_PROFILING.md).
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
QUICK="--steps 1 --warmup 3 --no-cpu-baseline --no-profile-pass --no-library-baseline"
python bench.py --dump-kernels $O/${TAG}_gemm_launches.txt > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_reference_arm.err
python bench.py --workload infer --steps 5 > $O/${TAG}_bench_infer.json 2> $O/${TAG}_bench_infer.err
python bench.py --workload cv --steps 5 > $O/${TAG}_bench_cv.json 2> $O/${TAG}_bench_cv.err
python bench.py $QUICK > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py $QUICK > $O/${TAG}_ncu_list.log 2>&1
# one full capture per GEMM kernel; launch indices pick big layers of the second (eager) step:
# igemm_pair_kernel 28 launches per step (the 8^3 level runs on it in split-K form): #13-#15 of the second step = up3.conv1 fprop (928 GF, 256 -> 128 at 64^3),
# up3.conv2 fprop (464 GF), up4.conv1 dgrad (1855 GF, 64 -> 128 at 128^3).  dmarch2_kernel 6 per step (#1 inc.conv2
# fprop 928 GF, #2 up4.conv1 fprop 1855 GF), wgrad_halo_kernel 15 per step (#1 up4.conv2 64 -> 64, #2 up4.conv1),
# conv1_march_kernel / conv1_march_wgrad_kernel one per step (the first layer).
for spec in "igemm_pair_kernel:40:3:igemm_pair" "dmarch2_kernel:6:2:dmarch2" "wgrad_halo_kernel:15:2:wgrad_halo" \
            "conv1_march_kernel:1:1:conv1_march" "conv1_march_wgrad_kernel:1:1:conv1_march_wgrad"; do
    IFS=: read -r kern skip cnt name <<< "$spec"
    python bench.py $QUICK > $O/${TAG}_plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on --kernel-name $kern --launch-skip $skip --launch-count $cnt \
        -f -o $O/${TAG}_prof_${name} python bench.py $QUICK > $O/${TAG}_ncu_${name}.log 2>&1
done
# the HBM-bound BatchNorm / pooling passes at the full-resolution shape (2 x 128^3 x 64): one launch each
python tools/bench_fused_bn_l0.py > $O/${TAG}_plain_bn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"bn_bwd|bn_apply|maxpool" -s 9 -c 9 \
    -f -o $O/${TAG}_prof_bn_passes python tools/bench_fused_bn_l0.py > $O/${TAG}_ncu_bn_passes.log 2>&1
python tools/bench_fused_bn.py > $O/${TAG}_bn_passes_timing.txt 2>&1
python tools/ab_fused.py > $O/${TAG}_ab_fused.txt 2>&1
python tools/ab_pdl.py > $O/${TAG}_ab_pdl.txt 2>&1
python tools/bench_dmarch2.py > $O/${TAG}_dmarch2_bench.txt 2>&1
python tools/bench_conv1_march.py > $O/${TAG}_conv1_march_bench.txt 2>&1
ls -la $O | tail -20
