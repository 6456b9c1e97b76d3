#!/bin/bash
# rebuild the CUDA library in-tree (sm_100a) from any cwd
cd "$(dirname "$0")/.." && python -c "from sonar_3d_reconstruction_b200.build import build_native; print(build_native(force=True))" 2>&1 | grep -v "^$"
