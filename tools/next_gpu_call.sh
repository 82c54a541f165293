#!/usr/bin/env bash
# First GPU calls of the next round (what round 2 could not re-measure after its GPU budget was spent):
#   gpurun --timeout 900 -- 'bash tools/run_evidence_1gpu.sh'          pytest -m gpu, bench line with both baselines,
#                                                                       per-launch table, timeline (one B200)
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/run_peer_8gpu.sh'      in-switch exchange at 8 GPUs: kernel check,
#                                                                       trainer check, iteration bench (in-switch / unicast)
#   gpurun --gpus 4 --timeout 600 -- 'bash tools/run_dp_4gpu.sh'        4 GPUs: default / in-switch / sharded step
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/run_sharded_2gpu.sh'   2 GPUs: dp_check (default, sharded), benches
# Open items with a measured starting point: narrow-layer tap stacking, Adam + re-pack fusion,
# bucketed exchange of the generator's gradient, sharded step with the pull and push overlapped.
set -u
bash tools/run_evidence_1gpu.sh
