mkdir -p gpurun_out; out=gpurun_out/x10_knobs.log; : > $out
for b in 20 24 26 28 30; do for f in 3 4 5 6; do echo -n "BATCH=$b FRAC8=$f " >> $out; B200RT_BATCH=$b B200RT_FRAC8=$f timeout 60 python tools/profile_frame.py 250 2>&1 | tail -1 >> $out; done; done
for v in S2 S4 S5; do echo -n "variant $v " >> $out; B200RT_LIB=$PWD/build/variants/libb200rt_$v.so timeout 60 python tools/profile_frame.py 250 2>&1 | tail -1 >> $out; done
for f in 3 4 6; do echo -n "variant S4 FRAC8=$f " >> $out; B200RT_FRAC8=$f B200RT_LIB=$PWD/build/variants/libb200rt_S4.so timeout 60 python tools/profile_frame.py 250 2>&1 | tail -1 >> $out; done
cat $out
