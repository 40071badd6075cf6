import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))
import rtdm_b200 as rt
l = rt.lib()
out = (C.c_double * 15)()
l.rtdm_dev_op_rates.restype = C.c_int
rc = l.rtdm_dev_op_rates(0, out, 15)
names = ["IADD3", "VIMNMX.U16x2+LOP3", "VABSDIFF4+IADD", "PRMT", "SHF", "VIADD.16x2", "IMAD", "VIMNMX3.U16x2", "VIMNMX.U32", "LOP3(x2?)", "VIMNMX3.U32", "VIMNMX.U16x2", "VABSDIFF4", "VSUB2", "VIADDMNMX.S16x2"]
print("rc", rc)
for n, v in zip(names, out):
    print(f"{n:20s} {v:7.2f} T stmt/s  -> {v * 1e12 / (148 * 1.965e9):6.1f} lanes/clk/SM @1965MHz")
