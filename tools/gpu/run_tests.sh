#!/bin/bash
# pytest -m gpu on the box; logs to gpurun_out/
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-t}
shift
timeout 900 python -m pytest tests -m gpu -q --timeout 300 "$@" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"
tail -25 gpurun_out/${tag}_pytest.log
