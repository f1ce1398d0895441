#!/usr/bin/env python3
"""Commit the SASS of the hot kernels (BASELINE north_star: "with the SASS committed"): tools/dump_sass.py writes profiles/sass/<kernel>.sass
— the instruction stream of each kernel (encodings stripped) — and profiles/sass/mnemonics.txt with the Blackwell-native mnemonic counts
(UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, FFMA2 = packed fp32 FMA, SYNCS = mbarrier)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJS = [os.path.join(ROOT, 'tf-keras-deeplabv3p-model-set_b200', 'build', f) for f in ('dlv3p_api.o', 'xception_api.o')]
HOT = [  # (file stem, substrings the demangled name must contain)
    ('bb_gemm_kernel_192', ['bb_gemm_kernel<(int)192, (bool)0>']),
    ('bb_gemm_kernel_256_residual', ['bb_gemm_kernel<(int)256, (bool)1>']),
    ('bb_depthwise_kernel_s1_r1', ['bb_depthwise_kernel<(int)1, (int)1, (int)8, (int)32>']),
    ('bb_sepconv_kernel_kb2_n128', ['bb_sepconv_kernel<(int)2, (int)128']),
    ('conv3x3_c32_kernel', ['conv3x3_c32_kernel']),
    ('stem_tc_kernel', ['stem_tc_kernel']),
    ('stem_conv_kernel', ['stem_conv_kernel<(bool)0>']),
    ('pw_gemm2_kernel', ['dlv3p::pw_gemm2_kernel']),
    ('dwpw_gemm2_kernel_kb5', ['dwpw_gemm2_kernel<(int)5>']),
    ('aspp_dw_fast3_kernel_32', ['aspp_dw_fast3_kernel<(int)32, (int)32, (int)6>']),
    ('resize_argmax_x4_kernel', ['resize_argmax_x4_kernel']),
    ('tgemm_kernel_256_tn', ['tgemm_kernel<(int)256, (bool)1>']),
    ('p2p_allreduce_kernel', ['p2p_allreduce_kernel']),
]
MNEMONICS = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTMALDG', 'UTMASTG', 'LDTM', 'UTCBAR', 'SYNCS', 'FFMA2', 'HMNMX2', 'LDS', 'STS', 'LDG', 'STG', 'UCGABAR_ARV']


def main():
    out_dir = os.path.join(ROOT, 'profiles', 'sass')
    os.makedirs(out_dir, exist_ok=True)
    funcs = {}
    for obj in OBJS:
        txt = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
        mangled = re.findall(r'Function : (\S+)', txt)
        dem = subprocess.run(['cu++filt'] + mangled, capture_output=True, text=True).stdout.splitlines() if mangled else []
        parts = re.split(r'\n\s*Function : \S+\n', txt)[1:]
        for mg, dm, body in zip(mangled, dem, parts):
            funcs[dm] = (mg, body)
    table = []
    for stem, pats in HOT:
        hit = [k for k in funcs if all(p in k for p in pats)]
        if not hit:
            print('not found:', stem, file=sys.stderr)
            continue
        name = hit[0]
        mg, body = funcs[name]
        lines = []
        for ln in body.splitlines():
            m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
            if m:
                lines.append('/*%s*/ %s ;' % (m.group(1), m.group(2).rstrip()))
        with open(os.path.join(out_dir, stem + '.sass'), 'w') as f:
            f.write('// %s\n// %s\n// sm_100a SASS, cuobjdump -sass (encodings stripped), %d instructions\n' % (name, mg, len(lines)))
            f.write('\n'.join(lines) + '\n')
        ops = [re.sub(r'^@!?U?P\d+\s+', '', ln.split('*/ ', 1)[1]).split()[0] for ln in lines]
        row = {mn: sum(1 for o in ops if (o == mn or o.startswith(mn + '.')) and not (mn == 'UTCHMMA' and '.2CTA' in o)) for mn in MNEMONICS}
        row['UTCHMMA.2CTA'] = sum(1 for o in ops if o.startswith('UTCHMMA') and '.2CTA' in o)
        table.append((stem, len(lines), row))
    with open(os.path.join(out_dir, 'mnemonics.txt'), 'w') as f:
        f.write('%-34s %6s ' % ('kernel', 'instr') + ' '.join('%12s' % m for m in MNEMONICS) + '\n')
        for stem, n, row in table:
            f.write('%-34s %6d ' % (stem, n) + ' '.join('%12d' % row[m] for m in MNEMONICS) + '\n')
    print(open(os.path.join(out_dir, 'mnemonics.txt')).read())


if __name__ == '__main__':
    main()
