"""Prints registers / spills per kernel from the *.ptxas.log files (dev tool)."""
import glob, re, subprocess, sys
for f in sorted(glob.glob(sys.argv[1] if len(sys.argv) > 1 else '*.ptxas.log')):
    txt = open(f).read()
    for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", txt):
        name = subprocess.run(['cu++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(.*', '', name.replace('b2q::', '').replace('void ', ''))
        print(f"{int(m.group(5)):4d} regs  stack {m.group(2):>5}  spill {m.group(3):>5}  {name}")
