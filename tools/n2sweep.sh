# 2..N-GPU bench sweep: exchange strategy (run under gpurun --gpus N)
N=${1:-2}
run() { # $1 = label, rest = bench args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 "${@:2}" 2> gpurun_out/n${N}_$1.err | grep '^{' | tee gpurun_out/n${N}_$1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['detail']['exchange'], 'slices', d['detail']['slices'], 'ms/step %.2f'%d['ms_per_step'], 'value %.3g'%d['value'], 'sweep ms %.3f'%d['roofline']['ms_per_sweep'], 'e2e ms %.1f'%d['e2e']['ms_per_step'], 'same', d['detail']['device_vs_host_arm_identical'])"
}
run p2p --comm p2p
run nccl --comm nccl
