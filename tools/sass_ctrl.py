#!/usr/bin/env python3
"""Decodes the control codes (static stall count, scoreboard barrier written, barriers waited for) of one symbol of the
clock-recovery core loop from the SASS of an object / cubin, and totals them: how many cycles the compiler's own schedule
costs, and how many shared-memory round trips are exposed.  No GPU needed.
usage: python tools/sass_ctrl.py build/obj/gr_cuda.o [v] [mangled kernel name]      (v = list every instruction)
default kernel: the time shards' clock-recovery kernel, mm_quad_kernel<SSE, 64>"""
import re,sys,subprocess
fun=[a for a in sys.argv[2:] if a.startswith('_Z')]
fun=fun[0] if fun else '_ZN3grb14mm_quad_kernelILi1ELi64EEEvNS_6MMArgsE'
out=subprocess.run(['cuobjdump','-sass','-fun',fun,sys.argv[1]],capture_output=True,text=True).stdout
lines=out.splitlines()
ins=[]; i=0
pat=re.compile(r'^\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/'); pat2=re.compile(r'^\s+/\* (0x[0-9a-f]{16}) \*/')
while i<len(lines):
    m=pat.match(lines[i])
    if m and i+1<len(lines):
        m2=pat2.match(lines[i+1])
        if m2:
            word=(int(m2.group(1),16)<<64)|int(m.group(3),16)
            ctrl=(word>>105)&((1<<21)-1)
            ins.append((int(m.group(1),16), m.group(2).strip(), ctrl&0xf,(ctrl>>5)&7,(ctrl>>11)&0x3f)); i+=2; continue
    i+=1
rm=[k for k,x in enumerate(ins) if 'FADD.RM' in x[1]]
a,b=rm[1],rm[2]
seg=ins[a:b]
print(sys.argv[1],'instr',len(seg),'static stalls',sum(x[2] for x in seg),'scoreboard waits',sum(1 for x in seg if x[4]), 'LDS',sum(1 for x in seg if x[1].startswith('LDS')))
if 'v' in sys.argv[2:]:
    for x in seg: print(hex(x[0]),x[1][:50].ljust(50),x[2],'wb',x[3] if x[3]!=7 else '-','wait',format(x[4],'06b'))
