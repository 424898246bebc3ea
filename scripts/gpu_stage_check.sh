#!/bin/bash
# Round-2 opener: validate and time the attention variants staged at the end of round 1 (off by default).
# Each block: parity tests of the attention kernels with the variant on (short timeouts: a barrier bug hangs), then timings.
mkdir -p gpurun_out
run() {  # $1 = tag, rest = env assignments
  tag=$1; shift
  env "$@" timeout 150 python -m pytest tests/test_gpu_kernels.py -k "attention" -x -q -p no:cacheprovider --timeout 30 \
      > gpurun_out/stage_$tag.log 2>&1
  echo "[$tag] pytest rc=$? $(tail -n 1 gpurun_out/stage_$tag.log)"
  if grep -q passed gpurun_out/stage_$tag.log && ! grep -q failed gpurun_out/stage_$tag.log; then
    env "$@" timeout 60 python scripts/prof_attn.py 2>&1 | grep -E "fwd .* bwd|attn_" | sed "s/^/[$tag] /"
    env "$@" PB=32 timeout 60 python scripts/prof_attn.py 2>&1 | grep -E "fwd .* bwd|attn_" | sed "s/^/[$tag b32] /"
  fi
}
run base LGB200_NOOP=1
run fwd_v3 LGB200_ATTN_FWD_V3=1
run dq_v4 LGB200_ATTN_DQ_V4=1
