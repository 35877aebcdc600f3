#!/bin/bash
# usage: run_variants.sh <workload> <genes> — EM-only timing for every built variant, init sweep on the default
cd /root/repo
for so in stminer_b200/csrc/variants/*.so; do
  STM_LIB_PATH=$PWD/$so timeout 120 python scripts/exp_em.py $1 $2 em 2>&1 | grep -v Warning
done
timeout 200 python scripts/exp_em.py $1 $2 init 2>&1 | grep -v Warning
