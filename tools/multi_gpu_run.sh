set -x
mkdir -p gpurun_out
nvidia-smi topo -m | head -12 > gpurun_out/topo8.txt; nproc >> gpurun_out/topo8.txt; lscpu | grep -E "NUMA|Socket" >> gpurun_out/topo8.txt
timeout 300 python -m pytest tests/test_ddp.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/nccl_test.log
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 100 --warmup 5 2>/dev/null | tail -1 > gpurun_out/scale_$n.json
done
timeout 200 python bench.py --gpus 1 --steps 100 --warmup 5 --skip-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/scale_1.json
for n in 1 2 8; do
  if [ $n = 1 ]; then timeout 200 python bench.py --ddp --steps 30 2>/dev/null | tail -1 > gpurun_out/ddp_$n.json; else
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --ddp --gpus $n --steps 30 2>/dev/null | tail -1 > gpurun_out/ddp_$n.json; fi
done
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.load(open("gpurun_out/scale_%d.json"%n)); print(n, round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["h2d_gbs_per_rank"],1), "born", round(d["e2e_device_born"]["value"]), round(d["e2e_device_born"]["h2d_gbs_per_rank"],1))
    except Exception as e: print(n, "ERR", e)
for n in (1,2,8):
    try:
        d=json.load(open("gpurun_out/ddp_%d.json"%n)); print("ddp", n, round(d["value"]), round(d["ms_per_step"],3), "hot", round(d["hot_path_ms"],3), round(d["hot_path_frac"],3), "exposed", round(d["allreduce_exposed_ms"],3))
    except Exception as e: print("ddp", n, "ERR", e)
PY
cat gpurun_out/nccl_test.log gpurun_out/topo8.txt
