#!/bin/bash
# parity tests under the most aggressive neighbour-build variant, then the short bench under each variant
set -u
mkdir -p gpurun_out
TAG=${1:-nv}
( XNB_NBH_VARIANT=3 timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_v3.log 2>&1; tail -1 gpurun_out/${TAG}_pytest_v3.log
SKIP_TESTS=1 bash scripts/gpu_variants.sh ${TAG} "XNB_NBH_VARIANT=0" "XNB_NBH_VARIANT=1" "XNB_NBH_VARIANT=2" "XNB_NBH_VARIANT=3" "XNB_NBH_VARIANT=0"
