import ctypes as C, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lfr_b200 import build_problem, synth
from lfr_b200.capi import load_b200
lib=load_b200()
for cfg in sys.argv[1:] or ['cfg2']:
    p=build_problem(synth.generate(cfg))
    s,arrays=lib.marshal(p)
    us=C.c_double(); nl=C.c_int()
    f=lib.lib.lfr_debug_time_schedule; f.argtypes=[C.c_void_p,C.c_void_p,C.c_int,C.POINTER(C.c_double),C.POINTER(C.c_int),C.c_void_p]
    rc=f(C.byref(s),None,50,C.byref(us),C.byref(nl),None); print(cfg,'rc',rc,'schedule us',us.value,'launches',nl.value)
