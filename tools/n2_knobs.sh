# N = 2 sweep of the knobs that could hide the gradient all-reduce of CapturedTrainStep (run under `gpurun --gpus 2`):
#   CCX_NCCL_HIGH_PRIO=1     NCCL kernels on a high-priority stream (their CTAs are placed before queued compute CTAs)
#   NCCL_CGA_CLUSTER_SIZE=1  no thread-block clusters for the NCCL kernels (no need for 4 free SMs in one GPC)
# usage: bash tools/n2_knobs.sh <tag> VAR=VALUE ...
tag=$1; shift
env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/n2_$tag.log 2> gpurun_out/n2_$tag.err
echo "$tag rc=$?"
python - <<PY
import json
for line in open("gpurun_out/n2_$tag.log"):
    if line.startswith("{"):
        d = json.loads(line)
        print("$tag", round(d["value"], 1), "img/s", round(d["ms_per_step"], 3), "ms", json.dumps(d.get("allreduce"))[:400])
PY
