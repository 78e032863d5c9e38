"""Per-kernel SASS comparison of two builds of libsdimb.so (instruction text, addresses ignored): python tools/sass_compare.py a.so b.so"""
import sys,re,subprocess
def funcs(lib):
    out=subprocess.run(["cuobjdump","-sass",lib],capture_output=True,text=True).stdout
    res={};name=None
    for line in out.splitlines():
        m=re.search(r"Function : (\S+)",line)
        if m: name=re.sub(r"_GLOBAL__N__[0-9a-f_]+sdimb_cu_[0-9a-f]+","NS",m.group(1)); res[name]=[]; continue
        m=re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);",line)
        if m and name: res[name].append(m.group(1))
    return res
a=funcs(sys.argv[1]); b=funcs(sys.argv[2])
for k in sorted(a):
    same = a[k]==b.get(k)
    print(len(a[k]), len(b.get(k,[])), "SAME" if same else "DIFF", k[:110])
