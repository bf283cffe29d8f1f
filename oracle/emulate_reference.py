"""TEST INFRASTRUCTURE.  Prepare the UNMODIFIED reference for the CUDA emulation of tests/emu, so that its own
kernels and solvers can be executed on a machine without a GPU (golden vectors for the parts of the oracle that
could not be pinned on a GPU box yet).

g++ cannot parse CUDA's launch syntax, so the reference's headers (read where they lie, /root/reference/include)
and the harness sources of this directory are mirrored into a scratch directory OUTSIDE the repository (default
/tmp/gv_b200_ref_emu, see oracle/Makefile) with two textual changes:

    kernel<<<grid, block[, shared[, stream]]>>>(args)   ->   gv_emu::LaunchConfig(grid, block, ...)(kernel)(args)
    model.backward<optimizer_type>(...)                 ->   model.template backward<optimizer_type>(...)

(the second is the `template` disambiguator ISO C++ requires for a dependent member template; nvcc's front end
accepts its absence, g++ does not).  Nothing else is touched; the mirror is a build intermediate like an nvcc --keep
directory: it never enters the repository (not even as an ignored file), only the shared objects built from it
land in oracle/_ref/.  oracle/Makefile (target ref_emu) compiles it with tests/emu/cuda_emu.h
force-included and cuRAND's device generator redirected to its host generator (same XORWOW stream; in the
emulation "device" memory is host memory).
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LAUNCH = re.compile(r"((?:[A-Za-z_]\w*::)*[A-Za-z_]\w*(?:<[^<>;(){}]*(?:<[^<>]*>[^<>;(){}]*)*>)?)\s*<<<([^;{}<>]*?)>>>", re.S)
DEPENDENT = re.compile(r"\bmodel\.backward<")


def rewrite(text):
    text = LAUNCH.sub(lambda m: "gv_emu::LaunchConfig(%s)(%s)" % (m.group(2).strip(), m.group(1)), text)
    return DEPENDENT.sub("model.template backward<", text)


def mirror(source, destination):
    count = 0
    for directory, _, files in os.walk(source):
        for name in files:
            if not name.endswith((".h", ".cuh", ".cu", ".hpp")):
                continue
            path = os.path.join(directory, name)
            target = os.path.join(destination, os.path.relpath(path, source))
            os.makedirs(os.path.dirname(target), exist_ok=True)
            with open(path, "r", errors="replace") as fin:
                text = fin.read()
            new = rewrite(text)
            count += len(LAUNCH.findall(text))
            with open(target, "w") as fout:
                fout.write(new)
    return count


def main(reference, out):
    launches = mirror(os.path.join(reference, "include"), os.path.join(out, "include"))
    os.makedirs(os.path.join(out, "src"), exist_ok=True)
    for name in ("ref_harness.cu", "ref_harness_kg.cu"):
        with open(os.path.join(HERE, name)) as fin:
            text = fin.read()
        launches += len(LAUNCH.findall(text))
        with open(os.path.join(out, "src", name), "w") as fout:
            fout.write(rewrite(text))
    print("mirrored the reference into %s: %d kernel launches rewritten" % (out, launches))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference",
         sys.argv[2] if len(sys.argv) > 2 else "/tmp/gv_b200_ref_emu")
