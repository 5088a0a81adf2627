for c in hnerv_l nerv_s enerv_m; do
  timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:bnerv --csv --log-file gpurun_out/r02_frame_traffic_$c.csv python tools/frame_once.py $c 4 > /dev/null 2>&1
done
ls -la gpurun_out | tail -5
