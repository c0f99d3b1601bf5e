mkdir -p gpurun_out/r02f
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02f/gpu_tests.log 2>&1; tail -3 gpurun_out/r02f/gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02f/smoke.log 2>&1; tail -1 gpurun_out/r02f/smoke.log
python - > gpurun_out/r02f/three_crops.txt 2>&1 <<'PY'
import sys; sys.path.insert(0, "fake-video-detection-engine_b200")
import torch, v5ela
from v5ela.batch import get_handle
from v5ela.synth import gen_frame
crops = [torch.from_numpy(gen_frame(i, 257, 301, 1)).cuda() for i in range(3)]
hd = get_handle(0)
for stage in ("smem", "mma"):
    hd.block_stage = stage
    for _ in range(3): v5ela.analyze_ragged(crops)
    torch.cuda.synchronize(); hd.profile_enable(True); hd.profile_read(True)
    for _ in range(50): v5ela.analyze_ragged(crops)
    ms, cnt = hd.profile_read(True); hd.profile_enable(False)
    print(f"three 257x301 crops, ragged call, block stage {stage}: {ms / cnt * 1e3:.1f} us of fused-kernel time per call")
PY
cat gpurun_out/r02f/three_crops.txt
timeout 600 python bench.py > gpurun_out/r02f/bench_c2.json 2> gpurun_out/r02f/bench_c2.err
for c in 1 3 4 5; do timeout 600 python bench.py --config $c --steps 3 --no-cpu > gpurun_out/r02f/bench_c$c.json 2> gpurun_out/r02f/bench_c$c.err; done
timeout 600 python bench.py --block-stage mma --no-cpu --no-files > gpurun_out/r02f/bench_c2_mma.json 2> gpurun_out/r02f/bench_c2_mma.err
if [ -z "$V5_FINAL_SHORT" ]; then
timeout 600 python bench.py --impl reference --steps 3 > gpurun_out/r02f/bench_ref.json 2> gpurun_out/r02f/bench_ref.err
python profiles/h2d_probe.py > gpurun_out/r02f/h2d_probe_1gpu.json 2>&1
python profiles/spectrum_perf.py > gpurun_out/r02f/spectrum.txt 2>&1
fi
python bench.py --steps 2 --warmup 1 --min-seconds 0 --no-cpu --no-parity > gpurun_out/r02f/plain.log 2>&1 && ncu --kernel-name-base demangled -k regex:v5 --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f/launches.csv python bench.py --steps 2 --warmup 1 --min-seconds 0 --no-cpu --no-parity > gpurun_out/r02f/ncu_launches.log 2>&1
python profiles/ncu_fused_case.py > gpurun_out/r02f/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ela_fused -s 2 -c 1 -o gpurun_out/r02f/fused_smem -f python profiles/ncu_fused_case.py > gpurun_out/r02f/ncu_smem.log 2>&1
V5ELA_BLOCK_STAGE=mma python profiles/ncu_fused_case.py > gpurun_out/r02f/ncu_plain2.log 2>&1 && V5ELA_BLOCK_STAGE=mma ncu --set full --clock-control none --import-source on -k regex:ela_fused -s 2 -c 1 -o gpurun_out/r02f/fused_mma -f python profiles/ncu_fused_case.py > gpurun_out/r02f/ncu_mma.log 2>&1
ls -la gpurun_out/r02f/
