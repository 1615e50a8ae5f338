mkdir -p gpurun_out
for rep in 1 2; do for v in q2r q3r q4r; do
  SELENITE_B200_LIB=build/ab/lib$v.so timeout 200 python tools/bench_chains.py --which q15 --steps 10 > gpurun_out/s26_q15_${v}_$rep.json 2>&1; echo -n "$v rep $rep: "; tail -1 gpurun_out/s26_q15_${v}_$rep.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['Gsamples_per_s'],1), round(d['hbm_frac'],4))"
done; done
