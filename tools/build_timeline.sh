#!/bin/bash
cd "$(dirname "$0")/.."
python - <<'PY'
from spn4cir_b200 import _lib
print(_lib.build_library(force=True, extra_flags=["-DSPN_TIMELINE"], out=_lib.PKG_DIR / "libspn_b200_timeline.so"))
PY
