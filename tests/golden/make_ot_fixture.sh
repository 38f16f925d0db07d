#!/bin/bash
# Builds and runs make_ot_fixture.cpp against the libtorch inside the installed torch wheel (CPU only).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
T=$(python -c "import torch, os; print(os.path.dirname(torch.__file__))")
g++ -std=c++17 -O1 -D_GLIBCXX_USE_CXX11_ABI=1 -I"$T/include" -I"$T/include/torch/csrc/api/include" "$HERE/make_ot_fixture.cpp" \
    -o /tmp/make_ot_fixture -L"$T/lib" -ltorch -ltorch_cpu -lc10 -Wl,-rpath,"$T/lib"
/tmp/make_ot_fixture "$HERE/tch_varstore.ot"
ls -la "$HERE/tch_varstore.ot"
