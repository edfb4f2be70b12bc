set -x
cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
for v in "" k3b3 k3b2; do
  if [ -z "$v" ]; then unset MARXB200_LIB; else export MARXB200_LIB=/root/repo/build/variants/libmarxb200_$v.so; fi
  echo "=== variant [$v]"
  timeout 300 python tools/trace_probe.py 16777216 c2_hetg_acis_s 10 2>&1 | tail -1
done
unset MARXB200_LIB
timeout 300 python tools/trace_probe.py 16777216 c4_beta_acis_i 10 2>&1 | tail -1
timeout 300 python tools/trace_probe.py 16777216 c1_acis_s 10 2>&1 | tail -1
# drop-in driver end to end: 1e8 rays, HETG+ACIS-S (C2), 1e6-ray batches (marx.par's dNumRays maximum), output to tmpfs
export MARX_DATA_DIR=/root/repo/oracle/_ref/data
C2="ExposureTime=0 Verbose=0 SourceFlux=0.003 TStart=2023.5 SpectrumType=FLAT SourceType=POINT MinEnergy=0.3 MaxEnergy=8.0 GratingType=HETG DetectorType=ACIS-S DitherModel=INTERNAL RandomSeed=1"
rm -rf /dev/shm/mg /dev/shm/mg2 /dev/shm/mc
( time integration/_build/marx_gpu @@oracle/_ref/par/marx.par $C2 NumRays=100000000 dNumRays=1000000 OutputDir=/dev/shm/mg ) > gpurun_out/drv_bulk.log 2>&1; tail -12 gpurun_out/drv_bulk.log
du -sh /dev/shm/mg; ls /dev/shm/mg | wc -l
( export MARXB200_EGRESS=stock; time integration/_build/marx_gpu @@oracle/_ref/par/marx.par $C2 NumRays=100000000 dNumRays=1000000 OutputDir=/dev/shm/mg2 ) > gpurun_out/drv_stock.log 2>&1; tail -5 gpurun_out/drv_stock.log
cmp /dev/shm/mg/energy.dat /dev/shm/mg2/energy.dat && cmp /dev/shm/mg/time.dat /dev/shm/mg2/time.dat && echo IDENTICAL
( time oracle/_ref/marx @@oracle/_ref/par/marx.par $C2 NumRays=4000000 dNumRays=1000000 OutputDir=/dev/shm/mc ) > gpurun_out/drv_cpu.log 2>&1; tail -5 gpurun_out/drv_cpu.log
rm -rf /dev/shm/mg /dev/shm/mg2 /dev/shm/mc
