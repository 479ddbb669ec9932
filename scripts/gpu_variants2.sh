#!/bin/bash
V=oriented_object_detection_b200/lib/variants
python scripts/stage_times.py base
for v in "$@"; do
  GM_LIB_PATH=$V/$v.so python scripts/stage_times.py $v
done
