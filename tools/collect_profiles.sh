#!/bin/bash
# One gpurun call that refreshes the ncu evidence under gpurun_out/ (summarised into profiles/ by
# tools/summarize_ncu.py).  usage: bash tools/collect_profiles.sh [families...]   (default: all four)
fams="${@:-crop nms proposal semdist}"
for f in $fams; do
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02_launches_$f.csv python tools/prof_driver.py $f 1 > /dev/null 2>&1
done
for f in $fams; do
  case $f in
    crop) rx="crop_"; cnt=12;;
    nms) rx="nms_|rank_"; cnt=8;;
    semdist) rx="layer_|edt_"; cnt=4;;
    *) continue;;
  esac
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -o gpurun_out/r02_full_$f -f python tools/prof_driver.py $f 1 > /dev/null 2>&1
done
ls -la gpurun_out
