# development loop: frame times of build variants (build/variants/libb200rt_<X>.so; D = the in-tree default build)
mkdir -p gpurun_out; out=gpurun_out/${EXP_NAME:-x}_perf.log; : > $out
for v in ${EXP_VARIANTS:-D}; do
  if [ $v = D ]; then lib=a_dive_into_ray_tracing_b200/libb200rt.so; else lib=build/variants/libb200rt_$v.so; fi
  echo "== variant $v" >> $out
  B200RT_LIB=$PWD/$lib timeout 90 python tools/quick_perf.py ${EXP_CASES:-c2 c4} >> $out 2>&1
done
cat $out
