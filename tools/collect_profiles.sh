set -x
for f in crop nms proposal semdist; do
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r01_launches_$f.csv python tools/prof_driver.py $f 1 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"crop_" -c 10 -o gpurun_out/r01_full_crop -f python tools/prof_driver.py crop 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nms_|rank_" -c 8 -o gpurun_out/r01_full_nms -f python tools/prof_driver.py nms 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"layer_|edt_" -c 4 -o gpurun_out/r01_full_semdist -f python tools/prof_driver.py semdist 1 > /dev/null 2>&1
ls -la gpurun_out
