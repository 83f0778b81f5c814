#!/bin/bash
# one ncu --set full capture of a 1 Mi-row encode with a variant library
mkdir -p gpurun_out; rm -f gpurun_out/p10*
export HIDVAE_B200_LIB=$PWD/hid-vae_b200/build/variants/${PLIB:-s3}.so
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:rq_fwd_tc" -s 4 -c 1 -o gpurun_out/p10_enc -f python tools/profile_step.py --rows 1048576 --reps 2 > gpurun_out/p10_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/p10_ncu.log
