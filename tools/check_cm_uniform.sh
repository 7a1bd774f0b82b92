#!/bin/bash
# The constant-memory (CM) variants of fused4_kernel only pay off when the compiler reads the matrices with
# uniform constant loads (LDCU ... c[0x3][URx]); whether it does depends on fragile heuristics of nvcc/ptxas
# (see DESIGN.md section 5).  This lists, for every CM instantiation in the built library, how many of its
# constant-bank-3 loads are uniform and how many fell back to per-thread LDC.
lib=${1:-$(dirname "$0")/../phyly_b200/lib/libarbplf_b200.so}
cuobjdump -sass "$lib" | awk '
/Function : /{name=$3; cm=(name ~ /fused4_kernelILi[0-9]+ELi[0-9]ELi[0-9]+ELi[0-9]ELb[01]ELb1E/)}
cm && /c\[0x3\]\[UR/ {u[name]++}
cm && /c\[0x3\]\[R/ {v[name]++}
cm {seen[name]=1}
END {bad=0; for (n in seen) { printf "%-75s uniform %4d  per-thread %4d\n", n, u[n], v[n]; if (v[n] > 0) bad=1 } exit bad}'
