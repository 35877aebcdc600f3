"""Builds a tuning / profiling variant of the library next to the default one:
   python scripts/build_variant.py NAME DEFINE [DEFINE...]   ->  stminer_b200/csrc/libstminer_b200_NAME.so  (use via STM_LIB_PATH)"""
import os
import sys

sys.path.insert(0, ".")
from stminer_b200 import build

name, defines = sys.argv[1], sys.argv[2:]
out = os.path.join(build.CSRC, "libstminer_b200_%s.so" % name)
print(build.build(defines=defines, out=out))
