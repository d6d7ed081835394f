#!/bin/bash
# A/B of the MIC kernel's multiply forms on one box: rebuild features_mic.o with each flag set, time bench --workload mic.
cd sound-event-localization-detection_b200/csrc
for v in "-DMIC_SCALAR_CMUL -DMIC_SCALAR_PHAT" "-DMIC_SCALAR_CMUL" ""; do
  touch features_mic.cu
  make PTXASV="$v" > /dev/null 2>&1
  echo "== flags: [$v]"
  (cd ../.. && python bench.py --workload mic --clips 256 --steps 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])")
done
