"""Error budget of the long-horizon drift, mechanism by mechanism (DESIGN.md "Long-horizon accuracy").

    python tools/drift_budget/drift_budget.py [bits ...]

drift_budget.cpp restates the PHYSICS of one step (no reward) with a precision switch per quantity
(bit set = that quantity is carried / computed in fp64 instead of fp32):
    1 fuel mass carried        2 angle and angle delta carried     4 x and dx carried     8 y and dy carried
   16 all force arithmetic    32 total mass and 1/m                64 sin / cos
  128 the SHIPPED scheme: fuel / angle fixed point, 1/m and alpha dt^2 in fp64 (with bit 2: angle delta in fp64)
  256 dy only   512 y only   1024 dx, dy as 2^-23 m fixed point   2048 angle delta = fp32 + 8 bits
 4096 x, y as 2^-15 m fixed point (with 1024)
and the script flies 512 rockets for 1000 steps under the hover controller of tests/parity.py against
the fp64 C oracle, printing the worst error of each state word in units of the tolerance
1e-5 * max(|ref|, scale).  bits = 0 is the round-1 kernel's arithmetic; 7298 (= 4096 + 2048 + 1024 + 128 + 2)
is what csrc/rl_device.cuh does now.  TEST / ANALYSIS TOOL: not part of the product, uses the oracle."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
sys.path[:0] = [os.path.join(_ROOT, "tests"), _ROOT]
import parity                      # noqa: E402
from oracle import c_oracle        # noqa: E402

_SO = os.path.join(_HERE, "libdrift_budget.so")
if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "drift_budget.cpp")):
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", _SO,
                           os.path.join(_HERE, "drift_budget.cpp")])
L = C.CDLL(_SO)
L.abl_step.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
hover = parity.hover_actions
def run(bits,n=512,T=1000,seed=5,fn=hover):
    rng=np.random.default_rng(seed); P=parity.load_params()
    init=parity.sample_reset_states(rng,P,n); o=c_oracle.BatchOracle(n); o.inject(init)
    f32=lambda v:v.astype(np.float32).astype(np.float64)
    st=np.zeros((n,10)); st[:,0]=f32(init[:,0]);st[:,1]=f32(init[:,1]);st[:,2]=f32(init[:,6]);st[:,3]=f32(init[:,2]);st[:,4]=f32(init[:,3]);st[:,5]=f32(init[:,7]);st[:,6]=f32(init[:,8]);st[:,7]=f32(init[:,9]);st[:,9]=1
    # oracle starts from fp64 init; to isolate arithmetic use f32-rounded init for the oracle too
    init32=init.copy()
    for j in (0,1,2,3,6,7,8,9): init32[:,j]=f32(init[:,j])
    o.inject(init32)
    out=np.zeros((n,3)); SC=parity.SCALE; mx=np.zeros(8); alive=np.ones(n,bool)
    for t in range(T):
        a=fn(t,o,rng); L.abl_step(bits,st.ctypes.data,n,a.ctypes.data,out.ctypes.data)
        _,_,te,tr,_=o.step(a); ref=o.state()
        got=np.stack([st[:,0],st[:,1],out[:,0],out[:,1],st[:,2],out[:,2],st[:,7]],1)
        rf=np.stack([ref[:,0],ref[:,1],ref[:,2],ref[:,3],ref[:,6],ref[:,7],ref[:,9]],1)
        sc=np.array([SC[0],SC[1],SC[2],SC[3],SC[6],SC[7],SC[9]])
        err=np.abs(got-rf)/(1e-5*np.maximum(np.abs(rf),sc))
        alive&=~(te|tr)
        if alive.any(): mx[:7]=np.maximum(mx[:7],err[alive].max(0))
    return mx[:7], alive.sum()
if __name__=="__main__":
    names="x y vx vy ang om fuel".split()
    LABEL = {0: "round-1 arithmetic: everything fp32", 1: "+ fuel carried exactly", 2: "+ angle, angle delta carried exactly",
             4: "+ x, dx carried exactly", 8: "+ y, dy carried exactly", 16: "+ all forces in fp64", 32: "+ 1/m in fp64",
             64: "+ sin/cos in fp64", 3: "fuel + angle chain exact", 15: "every carried word exact, forces fp32",
             18: "angle chain exact + forces fp64", 31: "every carried word exact + forces fp64",
             130: "fuel / angle fixed point + fp64 torque chain (x, y, dx, dy still fp32)",
             386: "... + dy exact", 1154: "... + dx, dy 2^-23 m fixed point",
             3202: "... + angle delta fp32 + 8 bits (x, y still fp32)",
             7298: "SHIPPED scheme (rl_device.cuh): ... + x, y 2^-15 m fixed point"}
    for bits in [int(b) for b in sys.argv[1:]] or [0, 1, 2, 4, 8, 16, 32, 64, 3, 15, 18, 31, 130, 386, 1154, 3202, 7298]:
        mx, _ = run(bits)
        print(f"bits={bits:5d}  " + " ".join(f"{n}={e:6.3f}" for n, e in zip(names, mx)) + "   " + LABEL.get(bits, ""), flush=True)
