set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -s -k "cuda_graph or cross_request or teacache" 2>&1 | grep -E "passed|failed|rror|assert" | tail -6
