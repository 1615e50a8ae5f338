bash tools/profile_round.sh r02 > gpurun_out/r02_profile_round.log 2>&1; tail -3 gpurun_out/r02_profile_round.log | cut -c1-300
for w in tx chan q15; do bash tools/profile_chain.sh r02 $w 2 > /dev/null 2>&1; done
ls -la gpurun_out/r02_* | head -20
