#!/bin/bash
# Multi-GPU call: node-sharded parity + node-sharded and utterance-sharded bench lines at N = $1 GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > gpurun_out/c5_topo_n$N.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/c5_pytest_dist.log 2>&1; echo "rc=$?" >> gpurun_out/c5_pytest_dist.log
  timeout 300 $TR scripts/dist_check.py 4 2 > gpurun_out/c5_dist_check_n2.log 2>&1
fi
timeout 400 $TR bench.py --gpus $N --steps 30 --warmup 3 --no-cpu > gpurun_out/c5_bench_cfg2_n$N.json 2> gpurun_out/c5_bench_cfg2_n$N.err
case $N in
  2) timeout 400 $TR bench.py --gpus $N --workload cfg3 --shard nodes --steps 10 --warmup 3 > gpurun_out/c5_nodes_cfg3_n$N.json 2> gpurun_out/c5_nodes_cfg3_n$N.err ;;
  4) timeout 400 $TR bench.py --gpus $N --workload cfg3 --shard nodes --steps 10 --warmup 3 > gpurun_out/c5_nodes_cfg3_n$N.json 2> gpurun_out/c5_nodes_cfg3_n$N.err
     timeout 400 $TR bench.py --gpus $N --workload cfg3 --shard nodes --steps 10 --warmup 3 --chunks 1 > gpurun_out/c5_nodes_cfg3_n${N}_chunks1.json 2> gpurun_out/c5_nodes_cfg3_n${N}_chunks1.err
     timeout 400 $TR bench.py --gpus $N --workload cfg3 --shard nodes --steps 10 --warmup 3 --reserve-sms 0 > gpurun_out/c5_nodes_cfg3_n${N}_rs0.json 2> gpurun_out/c5_nodes_cfg3_n${N}_rs0.err
     timeout 400 $TR bench.py --gpus $N --workload cfg3 --shard nodes --steps 10 --warmup 3 --chunks 8 > gpurun_out/c5_nodes_cfg3_n${N}_chunks8.json 2> gpurun_out/c5_nodes_cfg3_n${N}_chunks8.err
     timeout 400 $TR bench.py --gpus $N --workload cfg3 --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/c5_bench_cfg3_n$N.json 2> gpurun_out/c5_bench_cfg3_n$N.err ;;
  8) timeout 400 $TR bench.py --gpus $N --workload cfg5 --shard nodes --steps 10 --warmup 3 --chunks 1 > gpurun_out/c5_nodes_cfg5_n${N}_chunks1.json 2> gpurun_out/c5_nodes_cfg5_n${N}_chunks1.err
     timeout 400 $TR bench.py --gpus $N --workload cfg5 --shard nodes --steps 10 --warmup 3 --chunks 2 --graph > gpurun_out/c5_nodes_cfg5_n${N}_graph2.json 2> gpurun_out/c5_nodes_cfg5_n${N}_graph2.err
     timeout 400 $TR bench.py --gpus $N --workload cfg5 --shard nodes --steps 10 --warmup 3 --chunks 4 --graph > gpurun_out/c5_nodes_cfg5_n${N}_graph4.json 2> gpurun_out/c5_nodes_cfg5_n${N}_graph4.err ;;
esac
ls -la gpurun_out | grep c5_ | head -20; for f in gpurun_out/c5_*n$N*.json; do head -c 400 $f; echo; done; tail -3 gpurun_out/c5_*n$N*.err 2>/dev/null | tail -20
