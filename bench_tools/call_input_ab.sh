#!/bin/bash
# A/B of the e2e input path on N GPUs (run on a GPU box, e.g. `gpurun --gpus N -- bash bench_tools/call_input_ab.sh N`):
#   feed+stream (default) | feed, plain stores | row gather (the path the round-2 e2e numbers were measured with)
# at the driver's invocation (K = 20) and in steady state (K = 2000). Results: gpurun_out/input_ab_nN/*.json + a summary.
N=${1:-1}
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/input_ab_n$N; mkdir -p $O
P=$((29000 + RANDOM % 300))
run() {   # name, steps, warmup, env...
  local name=$1 k=$2 w=$3; shift 3
  P=$((P+311))
  if [ "$N" = "1" ]; then
    env "$@" timeout 300 python bench.py --gpus 1 --steps $k --warmup $w --skip_parity > $O/${name}_k$k.json 2> $O/${name}_k$k.err
  else
    env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $N --steps $k --warmup $w --skip_parity > $O/${name}_k$k.json 2> $O/${name}_k$k.err
  fi
  echo "rc=$?" >> $O/${name}_k$k.err
}
for k in 20 2000; do
  w=5; [ $k = 2000 ] && w=50
  run feed_stream $k $w DM_EPOCH_FEED=1
  run feed_plain  $k $w DM_EPOCH_FEED=1 DM_STREAMING_COPY=0
  run gather      $k $w DM_EPOCH_FEED=0
  run gather_plain $k $w DM_EPOCH_FEED=0 DM_STREAMING_COPY=0
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        e = d.get("e2e", {})
        print(f.split("/")[-1], "N=$N value", round(d["value"]), "e2e", round(e.get("value", 0)), e.get("input_feed", {}).get("chunks_from_epoch_buffer"),
              e.get("input_feed", {}).get("chunks_row_gathered"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
