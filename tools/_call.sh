O=gpurun_out/c18
mkdir -p $O
PPS_CHEB_BLOCK=3 timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:cheb_blocked_kernel -s 40 -c 1 -f -o $O/prof_r02_256_cheb_blocked3 python tools/probe.py solve 256 cheb > $O/ncu_blocked.log 2>&1
PPS_CHEB_BLOCK=3 PPS_CHEB_F32=1 timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:cheb_blocked_kernel -s 40 -c 1 -f -o $O/prof_r02_256_cheb_blocked3_f32 python tools/probe.py solve 256 cheb > $O/ncu_blocked_f32.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:EpiChebStep -s 40 -c 1 -f -o $O/prof_r02_256_cheb_step python tools/probe.py solve 256 cheb > $O/ncu_step.log 2>&1
ls -la $O | head
