"""Integer-pipe microbenchmarks (p2v_int_pipe_peak): groups/s and groups per clock per SM."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plonky2_verifier_b200 as p2v
ctx = p2v.Context(0)
names = ["LOP3+IMAD.WIDE", "2LOP3+IMAD.WIDE", "LOP3+IMAD32", "LOP3+IADD(IMAD.IADD)", "2IMAD.WIDE+LOP3", "2xIMAD32", "2xLOP3", "IMAD.WIDE", "2xSHF", "IADD3+IADD3.X", "DFMA", "DFMA+LOP3+IMAD32", "DADD", "I2F.F64.U32+LOP3", "I2F+LOP3 | DFMA", "I2F+LOP3 | IMAD.WIDE", "IADD3 3reg", "LOP3 3reg", "DFMA 3reg", "DFMA 2reg+imm", "DFMA3 + IADD3", "DFMA2i + IADD3", "IMAD.WIDE acc", "IMAD.WIDE + DFMA2i + IADD3"]
out = {}
for m, nm in enumerate(names):
    v = ctx.int_pipe_peak(m)
    out[nm] = v
    print("%-24s %.3e groups/s  = %.1f groups/clk/SM (at 1965 MHz, 148 SMs)" % (nm, v, v / 148 / 1.965e9))
print(json.dumps(out))
