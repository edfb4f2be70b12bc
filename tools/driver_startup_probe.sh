#!/bin/bash
# developer probe: wall time of integration/_build/marx_gpu on a small C2 run (2^26 rays) with and without the early CUDA-context thread
cd "$(dirname "$0")/.."
export MARX_DATA_DIR=$PWD/oracle/_ref/data USER=${USER:-marx} MARXB200_TIMING=1
ARGS="ExposureTime=0 Verbose=0 SourceFlux=0.003 TStart=2023.5 SourceType=POINT SpectrumType=FLAT MinEnergy=0.3 MaxEnergy=8.0 GratingType=HETG DetectorType=ACIS-S DitherModel=INTERNAL"
for rep in 1 2 3; do for w in 1 0; do
  d=/dev/shm/mxb_startup_$$; rm -rf $d
  t0=$(date +%s.%N)
  MARXB200_WARMUP=$w integration/_build/marx_gpu @@integration/_build/par/marx.par OutputDir=$d NumRays=67108864 dNumRays=16777216 RandomSeed=1 $ARGS > /tmp/mxb_startup.log 2>&1
  t1=$(date +%s.%N)
  echo "warmup=$w wall $(python3 -c "print('%.3f' % ($t1 - $t0))") s  $(grep -o 'init+upload [0-9.]* (CUDA context [0-9.]*' /tmp/mxb_startup.log | tail -1)"
  rm -rf $d
done; done
