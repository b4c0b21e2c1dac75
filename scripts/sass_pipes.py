"""usage: sass_ops.sh lib fn | sass_pipes.py  -> instruction count per pipe class (ALU pipe = half rate on sm_100)."""
import sys, collections
def cls(op):
    b = op.split('.')[0]
    if b in ('FFMA', 'FMUL', 'FADD', 'IMAD', 'HFMA2', 'FFMA2'): return 'FMA'
    if b in ('PRMT', 'LOP3', 'SEL', 'FMNMX', 'FMNMX3', 'FSETP', 'ISETP', 'VIADD', 'SHF', 'IADD3', 'PLOP3', 'LEA', 'MOV', 'FSEL', 'VIMNMX', 'IABS', 'LOP', 'SGXT', 'BMSK', 'VIADDMNMX', 'VIMNMX3', 'R2P', 'P2R', 'IADD', 'FSET', 'CS2R'): return 'ALU'
    if b in ('POPC', 'FLO', 'BREV', 'MUFU', 'I2F', 'F2I', 'I2FP', 'FCHK'): return 'XU'
    if b in ('LDG', 'STG', 'LDS', 'STS', 'LDL', 'STL', 'ATOMG', 'RED', 'LDC', 'LDCU', 'ATOM', 'ATOMS'): return 'LSU'
    return 'other'
c = collections.Counter(); n = 0
for l in sys.stdin:
    t = l.split()
    if len(t) < 2: continue
    op = t[2] if t[1].startswith('@') else t[1]
    c[cls(op)] += 1; n += 1
print(n, dict(c))
