run() { # $1 = label, rest = env
  env "${@:2}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --slices ${SL:-4} 2> gpurun_out/n2_$1.err | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'slices', d['config']['slices'], 'ms/step %.2f'%d['ms_per_step'], 'value %.3g'%d['value'], 'e2e ms %.1f'%d['e2e']['ms_per_step'])"
}
SL=1 run base1 X=1
SL=4 run bps8_s4 HGE_BLOCKS_PER_SM=8
SL=4 run bps8_s4_hp HGE_BLOCKS_PER_SM=8 TORCH_NCCL_HIGH_PRIORITY=1
SL=4 run bps16_s4_hp HGE_BLOCKS_PER_SM=16 TORCH_NCCL_HIGH_PRIORITY=1
SL=2 run bps16_s2_hp HGE_BLOCKS_PER_SM=16 TORCH_NCCL_HIGH_PRIORITY=1
SL=8 run bps32_s8_hp HGE_BLOCKS_PER_SM=32 TORCH_NCCL_HIGH_PRIORITY=1
SL=1 run bps16_s1 HGE_BLOCKS_PER_SM=16
