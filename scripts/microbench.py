import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cam_nor_physics_b200 import soundings as S, zm_conv as Z
Z.zm_init(Z.default_params(16, 32, 3))
c = (C.c_longlong * 20)()
Z.lib().zm_microbench(c, 20, 200)
names = ["div", "log", "log10", "pow10", "exp", "es(T)", "enthalpy", "entropy", "ienthalpy", "ientropy", "pow",
         "goff-gratch formula", "dfma", "dadd", "dmul", "div_hot", "log_hot", "es(T) smem", "cmp+select", "f2i+i2f"]
print({n: int(v) for n, v in zip(names, c)})
