for v in "" nopf pf1 pf3; do
  if [ -n "$v" ]; then export GADDPG_LIB=/root/repo/ga-ddpg_b200/lib/libgaddpg_b200_$v.so; else unset GADDPG_LIB; fi
  echo "== variant ${v:-default(pf2)}"
  LEVELS=3 python scripts/bench_kernels2.py 2>&1 | grep -E "\(423608" | grep -v "tn dW"
  python scripts/bench_tn_pool.py 2>&1 | grep "nt dX"
done
