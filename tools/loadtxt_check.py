"""CPU-only check and timing of csrc/textio.cu's loader against np.loadtxt(..).astype('float32') (bit equality on every digit
count, float32 / double rounding midpoints, numpy's syntax) -- the long form of tests/test_abi.py::test_loadtxt_exact_parser_cases."""
import numpy as np, sys, os, time, struct
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfit_b200 import tutorial_utils as tu
rng=np.random.default_rng(3)
d='/tmp/ltx'; os.makedirs(d,exist_ok=True)
def check(path, note):
    ref=np.loadtxt(path).astype('float32')
    if ref.ndim==1: ref=ref.reshape(-1, ref.size if open(path).read().count('\n')<=1 else 1) if ref.ndim==1 else ref
    got=tu.loadtxt_f32(path)
    ref2=np.loadtxt(path, ndmin=2).astype('float32')
    ok = got.shape==ref2.shape and np.array_equal(got.view(np.uint32), ref2.view(np.uint32))
    print("%-52s %s %s"%(note, got.shape, "OK" if ok else "MISMATCH"))
    if not ok:
        bad=np.nonzero(got.view(np.uint32)!=ref2.view(np.uint32)) if got.shape==ref2.shape else None
        print("   ", bad, got[bad][:5] if bad else None, ref2[bad][:5] if bad else None)
    return ok
allok=True
# 1. every number of significant digits 1..25, random exponents, fixed and scientific notation
for nd in range(1,26):
    vals=rng.standard_normal((3000,3))*10.0**rng.integers(-12,12,size=(3000,3))
    p=os.path.join(d,'d%d.txt'%nd); np.savetxt(p, vals, fmt='%%.%de'%(nd-1)); allok&=check(p,"%%.%de scientific"%(nd-1))
for nd in (0,1,3,6,9,12,15,17,20):
    vals=rng.standard_normal((3000,3))*10.0**rng.integers(-3,6,size=(3000,3))
    p=os.path.join(d,'f%d.txt'%nd); np.savetxt(p, vals, fmt='%%.%df'%nd); allok&=check(p,"%%.%df fixed"%nd)
# 2. numpy's own format from float32 and float64 data, large file (multi-threaded path: > 1 MiB)
a=(rng.standard_normal((60000,3))*10.0**rng.integers(-8,8,size=(60000,3))).astype(np.float32)
p=os.path.join(d,'np32.txt'); np.savetxt(p,a); allok&=check(p,"np.savetxt of float32 (19 digits), 60000 rows")
a=(rng.standard_normal((60000,3))*10.0**rng.integers(-30,30,size=(60000,3)))
p=os.path.join(d,'np64.txt'); np.savetxt(p,a); allok&=check(p,"np.savetxt of float64 (19 digits), 60000 rows")
# 3. values at / next to float32 rounding midpoints written with 17-19 digits (double rounding traps)
f=rng.integers(0x30000000,0x50000000,size=20000,dtype=np.uint32).view(np.float32).astype(np.float64)
nxt=np.nextafter(f.astype(np.float32), np.float32(np.inf)).astype(np.float64)
mid=(f+nxt)/2
for k,fmt in enumerate(('%.16e','%.17e','%.18e')):
    p=os.path.join(d,'mid%d.txt'%k); np.savetxt(p, np.stack([mid, np.nextafter(mid,np.inf), np.nextafter(mid,-np.inf)],1), fmt=fmt); allok&=check(p,"float32 midpoints and their double neighbours "+fmt)
# 4. decimal strings at double midpoints: exact decimal expansion of (a+b)/2 truncated to 19 digits etc.
from decimal import Decimal, getcontext
getcontext().prec=60
rows=[]
for x in rng.standard_normal(4000)*10.0**rng.integers(-6,6,size=4000):
    x=float(x); y=float(np.nextafter(x,np.inf)); m=(Decimal(x)+Decimal(y))/2
    rows.append(" ".join(("%.18E"%m if False else format(m,'.18e'), format(m,'.17e'), format(m+Decimal(10)**(m.adjusted()-18),'.18e'))))
p=os.path.join(d,'dmid.txt'); open(p,'w').write("\n".join(rows)+"\n"); allok&=check(p,"decimal strings at double midpoints, 18-19 digits")
# 5. syntax: comments, blank lines, tabs, commas, CRLF, no trailing newline, signs, leading zeros, nan/inf, big exponents
txt="# header\n\n  1.5\t-2.25 ,+3e2 # tail\n\r\n0.000 -0.0 00012.50\r\n1e-320 1e308 1.7976931348623157e308\nnan inf -inf\n4.9e-324 2.2250738585072014e-308 123456789012345678901234567890\n.5 5. 1.e3"
p=os.path.join(d,'syn.txt'); open(p,'w').write(txt)
got=tu.loadtxt_f32(p); ref=np.array([[float(t) for t in l.split('#')[0].replace(',',' ').split()] for l in txt.replace('\r','').split('\n') if l.split('#')[0].strip()],dtype=np.float64).astype(np.float32)
ok=got.shape==ref.shape and np.array_equal(got.view(np.uint32),ref.view(np.uint32)); allok&=ok
print("syntax cases", got.shape, "OK" if ok else "MISMATCH", None if ok else (got, ref))
# 6. ragged row is an error
open(os.path.join(d,'rag.txt'),'w').write("1 2 3\n4 5\n")
try:
    tu.loadtxt_f32(os.path.join(d,'rag.txt')); print("ragged: NO ERROR"); allok=False
except Exception as e: print("ragged row ->", type(e).__name__)
# 7. empty file
open(os.path.join(d,'empty.txt'),'w').write("# nothing\n")
print("empty:", tu.loadtxt_f32(os.path.join(d,'empty.txt')).shape)
# 8. speed
a=rng.standard_normal((1000000,3)).astype(np.float32)
p=os.path.join(d,'big6.xyz'); np.savetxt(p,a,fmt='%.6f')
t=time.time(); g=tu.loadtxt_f32(p); t1=time.time()-t
t=time.time(); r=np.loadtxt(p,dtype=np.float64).astype(np.float32); t2=time.time()-t
print("1M x 3 '%%.6f' (%d MB): ours %.3f s (%.0f MB/s), numpy %.2f s, equal %s"%(os.path.getsize(p)/1e6,t1,os.path.getsize(p)/t1/1e6,t2,np.array_equal(g,r)))
p=os.path.join(d,'big18.xyz'); tu.savetxt(p,a)
t=time.time(); g=tu.loadtxt_f32(p); t1=time.time()-t
print("1M x 3 '%%.18e' (%d MB): ours %.3f s (%.0f MB/s), equal %s"%(os.path.getsize(p)/1e6,t1,os.path.getsize(p)/t1/1e6,np.array_equal(g,a)))
print("ALL OK" if allok else "FAILURES")
