"""Lists, per kernel of the built objects, the global-memory instructions that precede the kernel's
griddepcontrol.wait (SASS: ACQBULK) in program order.  A kernel launched programmatically must not touch anything its
predecessor wrote before that instruction; loads of `const __restrict__` data can be hoisted above it by the compiler
(pp_common.cuh, pdl_enter).  usage: python scripts/pdl_audit.py [objects...]"""
import glob, os, re, subprocess, sys
objs = sys.argv[1:] or sorted(glob.glob(os.path.join(os.path.dirname(__file__), "..", "objectdetection_3d_b200", "build", "*.o")))
for o in objs:
    out = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
    name, pre, has = None, [], False
    def flush():
        if name and has:
            print("%-14s %-60s %s" % (os.path.basename(o), name[:60], "; ".join(pre) if pre else "-"))
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            name, pre, has, seen = m.group(1), [], False, False
            nm = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", nm).replace("pp::(anonymous namespace)::", "").replace("void ", "")
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if not m or name is None:
            continue
        ins = m.group(1).strip()
        if "ACQBULK" in ins:
            has, seen = True, True
        elif not seen and re.search(r"\b(LDG|LD\.E|ATOMG|ATOM\.|RED\.|STG|ST\.E)", ins):
            pre.append(re.sub(r"\s+", " ", ins)[:48])
    flush()
