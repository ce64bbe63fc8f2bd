"""TEST INFRASTRUCTURE / integration proof: builds the REFERENCE's own slave with libsatsuma_b200 bound in.

    python oracle/make_refslave_b200.py            (or: make -C oracle refslave_b200)

Takes analysis/HomologyByXCorrSlave.cc where it lies under /root/reference and writes a patched translation unit
to oracle/_ref/obj/ (git-ignored build output, nothing of the reference is committed): three insertions, exactly
the edit INTEGRATION.md section 2 describes --
  1. satsuma2_b200/host/binding/slave_globals.inc  after the reference's includes,
  2. satsuma2_b200/host/binding/slave_setup.inc    in main() once targetTotal is known (Slave.cc:405-408),
  3. satsuma2_b200/host/binding/slave_align.inc    as the body of HomologyByXCorr::align_target (Slave.cc:270-300).
Everything else (flag parsing, FASTA loading, ChunkManager, worker threads, the TCP exchange with the master's
WorkQueue) stays the reference's code.  Output: oracle/_ref/HomologyByXCorrSlave_b200bind, linked against
satsuma2_b200/libsatsuma_b200.so.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("REF", "/root/reference")
BIND = os.path.join(ROOT, "satsuma2_b200", "host", "binding")
OUT_DIR = os.path.join(HERE, "_ref")
EXE = os.path.join(OUT_DIR, "HomologyByXCorrSlave_b200bind")


def patched_source() -> str:
    src = open(os.path.join(REF, "analysis", "HomologyByXCorrSlave.cc")).read()
    lines = src.split("\n")
    out, i, done = [], 0, set()
    while i < len(lines):
        ln = lines[i]
        if "globals" not in done and ln.startswith("#include <netdb.h>"):
            out += [ln, '#include "slave_globals.inc"']
            done.add("globals")
        elif "align" not in done and ln.startswith("void HomologyByXCorr::align_target(t_pair p)"):
            out += [ln, '#include "slave_align.inc"', "}"]
            depth = ln.count("{") - ln.count("}")
            while depth > 0:  # drop the reference body up to the matching brace
                i += 1
                depth += lines[i].count("{") - lines[i].count("}")
            done.add("align")
        elif "setup" not in done and "TIME SPENT ON LOADING" in ln:
            out += ['#include "slave_setup.inc"', ln]
            done.add("setup")
        else:
            out.append(ln)
        i += 1
    assert done == {"globals", "align", "setup"}, done
    return "\n".join(out)


def build() -> str:
    os.makedirs(os.path.join(OUT_DIR, "obj"), exist_ok=True)
    tu = os.path.join(OUT_DIR, "obj", "HomologyByXCorrSlave_b200bind.cc")
    with open(tu, "w") as f:
        f.write(patched_source())
    srcs = [tu] + [os.path.join(REF, p) for p in (
        "analysis/CrossCorr.cc", "analysis/DNAVector.cc", "analysis/AlignProbability.cc", "analysis/ProbTable.cc",
        "analysis/SeqChunk.cc", "analysis/CodonTranslate.cc", "analysis/SequenceMatch.cc", "analysis/WorkQueue.cc",
        "base/FileParser.cc", "base/StringUtil.cc", "util/mutil.cc", "util/SysTime.cc")]
    lib_dir = os.path.join(ROOT, "satsuma2_b200")
    cmd = ["g++", "-O3", "-w", "-std=c++14", "-pthread", "-include", "cstdint", "-include", "memory", "-I" + REF,
           "-I" + os.path.join(REF, "analysis"), "-I" + BIND, "-I" + os.path.join(ROOT, "include"), "-o", EXE] + srcs + [
           "-L" + lib_dir, "-lsatsuma_b200", "-Wl,-rpath," + lib_dir, "-Wl,-rpath,$ORIGIN/../../satsuma2_b200"]
    subprocess.run(cmd, check=True)
    return EXE


SHIM_EXE = os.path.join(OUT_DIR, "shim_check")


def build_shim_check() -> str:
    """oracle/shim_check.cc: the class-level shims (crosscorr_shim.h) side by side with the reference's own classes."""
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = [os.path.join(HERE, "shim_check.cc")] + [os.path.join(REF, p) for p in (
        "analysis/CrossCorr.cc", "analysis/DNAVector.cc", "analysis/CodonTranslate.cc", "base/FileParser.cc",
        "base/StringUtil.cc", "util/mutil.cc")]
    lib_dir = os.path.join(ROOT, "satsuma2_b200")
    cmd = ["g++", "-O2", "-w", "-std=c++14", "-pthread", "-include", "cstdint", "-include", "memory", "-I" + REF,
           "-I" + os.path.join(REF, "analysis"), "-I" + BIND, "-I" + os.path.join(ROOT, "include"), "-o", SHIM_EXE] + srcs + [
           "-L" + lib_dir, "-lsatsuma_b200", "-Wl,-rpath," + lib_dir, "-Wl,-rpath,$ORIGIN/../../satsuma2_b200"]
    subprocess.run(cmd, check=True)
    return SHIM_EXE


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not found: this target needs the reference sources")
    print(build())
    print(build_shim_check())
