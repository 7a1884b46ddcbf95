"""Static instruction mix of the kernels that bound the step, from the built library (no GPU): share of tensor / integer-address /
fp32 / SFU / memory / control instructions in the SASS, and the number of integer-division sequences (one IABS pair each).
Static counts say what the code is made of, not how often each line runs; they are read next to the dynamic numbers in
profiles/r01_ncu_final_metrics.csv.

    python tools/sass_mix.py > profiles/r01_static_instruction_mix.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tf_vqa_regat_b200", "libregat.so")
KERNELS = ["geoattn_fwd_bf16_kernelILi3", "attn_bwd_bf16_kernelILi3", "geo_bwd_kernelILi32ELb1", "gemm_tc_kernelILi256ELi4ELi2ELi8ELb0ELb0",
           "gemm_tc_kernelILi128ELi3ELi2ELi4ELb1ELb0", "gemm_tc_kernelILi64ELi8ELi2ELi4ELb0ELb0", "opt_update_kernel", "opt_reduce_kernel",
           "colsum_multi", "butd_pool_fwd_kernelI13", "butd_pool_bwd_kernelI13", "bce_kernelI13"]
GROUPS = [
    ("tensor", ("HMMA", "UTCHMMA", "UTCQMMA")),
    ("int / address / move", ("IMAD", "IADD", "LEA", "SHF", "LOP", "ISETP", "IMNMX", "VIMNMX", "SEL", "PRMT", "MOV", "UIMAD", "UIADD", "ULEA", "USHF",
                              "ULOP", "UMOV", "R2UR", "S2R", "S2UR", "CS2R", "UISETP", "USEL", "PLOP", "VIADD", "BMSK", "SGXT", "I2F", "F2I", "I2FP",
                              "IABS", "POPC", "FLO", "P2R", "R2P", "UPRMT", "UFLO", "UPOPC", "VABSDIFF")),
    ("fp32 / convert", ("FFMA", "FMUL", "FADD", "FMNMX", "FSETP", "FSEL", "F2F", "F2FP", "HADD", "HMUL", "HFMA", "FCHK", "FRND")),
    ("SFU", ("MUFU",)),
    ("memory", ("LDG", "STG", "LD", "ST", "RED", "ATOM", "LDS", "STS", "LDSM", "LDGSTS", "UTMA", "LDTM", "LDC", "ULDC", "UBLKCP", "UTMALDG")),
]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)[1:]
    names = subprocess.run(["c++filt"] + [b.split("\n", 1)[0].strip() for b in blocks], capture_output=True, text=True).stdout.split("\n")
    print("# Static instruction mix of the step's main kernels (tools/sass_mix.py, no GPU needed)\n")
    print("| kernel | SASS instr | " + " | ".join(g for g, _ in GROUPS) + " | sync / control / shuffle | integer divisions (IABS/2) |")
    print("|---|---|" + "---|" * (len(GROUPS) + 2))
    for b, dn in zip(blocks, names):
        raw = b.split("\n", 1)[0]
        if not any(k in raw for k in KERNELS):
            continue
        ops = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", b, flags=re.M)
        c = collections.Counter(ops)
        tot = sum(c.values())
        shares, used = [], 0
        for _, prefixes in GROUPS:
            n = sum(v for o, v in c.items() if o.startswith(prefixes))
            used += n
            shares.append(f"{100 * n / tot:.0f} %")
        dn = re.sub(r"regat::\(anonymous namespace\)::", "", re.sub(r"^void ", "", dn))
        dn = dn if len(dn) < 70 else dn[:67] + "..."
        print(f"| `{dn}` | {tot} | " + " | ".join(shares) + f" | {100 * (tot - used) / tot:.0f} % | {c.get('IABS', 0) // 2} |")
    print("\nReading: the bf16 attention kernels are more than half integer / address / move instructions around 84 `HMMA`s each; "
          "the forward kernel and the persistent GEMM also carry runtime integer divisions (last column: the `(i*M+j) div N` index "
          "scramble and row / head decompositions by runtime sizes; the GEMM's tile scheduler and epilogue addressing).  Both are "
          "candidates for compile-time sizes or multiply-shift division (DESIGN.md section 9, item 2).")


if __name__ == "__main__":
    main()
