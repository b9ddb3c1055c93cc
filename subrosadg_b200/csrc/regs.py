import re, sys
L = open('_build/sdg_api.ptxas.log').read().split('\n')
cur = None
for i, l in enumerate(L):
    m = re.search(r"Compiling entry function '(\S+)'", l)
    if m: cur = m.group(1)
    m = re.search(r"Used (\d+) registers", l)
    if m and cur and 'Stage' in cur:
        name = re.sub(r'_ZN3sdg\d+(\w+?)StageKernelILi(\d)ELi(\d)ELi(\d+)ELb(\d)ELi(\d).*', r'\1 D\2 N\3 K\4 aff\5 PH\6', cur)
        print(name, m.group(1), L[i - 1].strip()[:90])
