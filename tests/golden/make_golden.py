#!/usr/bin/env python
"""
Regenerate tests/golden/ from the reference checkout (run in the build
container only; /root/reference does not exist on the GPU box).

The fixtures are the reference's OWN golden vectors: every
examples/**/in*.json with its shipped out*.json (outputs of the Arb
implementation, correctly rounded doubles), plus the outputs that the
reference only quotes in README files.  Inputs are stored minified; outputs
are stored verbatim.  A manifest ties each input to its program and expected
output.

usage: python tests/golden/make_golden.py [/root/reference]
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

# (case name, program, input path under examples/, expected output path or inline dict)
CASES = [
    ("fels_ll", "ll", "Felsenstein.2004.fig.16.4/ll/in.json", "Felsenstein.2004.fig.16.4/ll/out.json"),
    # README.md:115-124 of Felsenstein.2004.fig.16.4/ll
    ("fels_ll2", "ll", "Felsenstein.2004.fig.16.4/ll/in2.json",
     {"columns": ["site", "value"],
      "data": [[0, 0.0], [1, -11.297288182875496], [2, -12.390132492111672]]}),
    ("fels_deriv", "deriv", "Felsenstein.2004.fig.16.4/deriv/in.json", "Felsenstein.2004.fig.16.4/deriv/out.json"),
    ("fels_marginal", "marginal", "Felsenstein.2004.fig.16.4/marginal/in.json", "Felsenstein.2004.fig.16.4/marginal/out.json"),
    ("fels_dwell_adenine", "dwell", "Felsenstein.2004.fig.16.4/dwell/adenine/in.json", "Felsenstein.2004.fig.16.4/dwell/adenine/out.json"),
    ("fels_dwell_pyrimidines", "dwell", "Felsenstein.2004.fig.16.4/dwell/pyrimidines/in.json", "Felsenstein.2004.fig.16.4/dwell/pyrimidines/out.json"),
    ("fels_trans_all", "trans", "Felsenstein.2004.fig.16.4/trans/all.types/in.json", "Felsenstein.2004.fig.16.4/trans/all.types/out.json"),
    ("fels_trans_AC", "trans", "Felsenstein.2004.fig.16.4/trans/A.to.C.only/in.json", "Felsenstein.2004.fig.16.4/trans/A.to.C.only/out.json"),
    ("fels_trans_tv", "trans", "Felsenstein.2004.fig.16.4/trans/transversions.only/in.json", "Felsenstein.2004.fig.16.4/trans/transversions.only/out.json"),
    ("fels_trans_tv2", "trans", "Felsenstein.2004.fig.16.4/trans/transversions.only/in2.json", "Felsenstein.2004.fig.16.4/trans/transversions.only/out2.json"),
    ("beast_jc69", "ll", "BEAST.JC69/in.json", "BEAST.JC69/out.json"),
    ("beast_k80", "ll", "BEAST.K80/in.json", "BEAST.K80/out.json"),
    ("beast_hky85", "ll", "BEAST.HKY85/in.json", "BEAST.HKY85/out.json"),
    ("beast_hky85g", "ll", "BEAST.HKY85G/in.json", "BEAST.HKY85G/out.json"),
    ("beast_hky85i", "ll", "BEAST.HKY85I/in.json", "BEAST.HKY85I/out.json"),
    ("beast_gtr", "ll", "BEAST.GTR/in.json", "BEAST.GTR/out.json"),
    ("beast_gtrg", "ll", "BEAST.GTRG/in.json", "BEAST.GTRG/out.json"),
    ("beast_gtri", "ll", "BEAST.GTRI/in.json", "BEAST.GTRI/out.json"),
    ("beast_gtrgi", "ll", "BEAST.GTRGI/in.json", "BEAST.GTRGI/out.json"),
    ("beast_anc_ll", "ll", "BEAST.AncestralState/ll/in.json", "BEAST.AncestralState/ll/out.json"),
    ("beast_anc_marginal", "marginal", "BEAST.AncestralState/marginal/in.json", "BEAST.AncestralState/marginal/out.json"),
    ("mj_jumps", "trans", "BEAST.MarkovJumps/MarkovJumpsC/in.json", "BEAST.MarkovJumps/MarkovJumpsC/out.json"),
    ("mj_jumps_reordered", "trans", "BEAST.MarkovJumps/MarkovJumpsC/in.reordered.json", "BEAST.MarkovJumps/MarkovJumpsC/out.reordered.json"),
    ("mj_rewards", "dwell", "BEAST.MarkovJumps/MarkovRewardsC/in.json", "BEAST.MarkovJumps/MarkovRewardsC/out.json"),
    ("mj_marginal_rate", "trans", "BEAST.MarkovJumps/MarkovMarginalRate/in.json", "BEAST.MarkovJumps/MarkovMarginalRate/out.json"),
    ("mj_marginal_rate_pedantic", "trans", "BEAST.MarkovJumps/MarkovMarginalRate/in.pedantic.json", "BEAST.MarkovJumps/MarkovMarginalRate/out.pedantic.json"),
    ("jc_long_ll", "ll", "JC.long.branch/ll/in.json", "JC.long.branch/ll/out.json"),
    ("jc_long_deriv", "deriv", "JC.long.branch/deriv/in.json", "JC.long.branch/deriv/out.json"),
    ("jc_long_marginal", "marginal", "JC.long.branch/marginal/in.json", "JC.long.branch/marginal/out.json"),
    ("jc_long_dwell", "dwell", "JC.long.branch/dwell/in.json", "JC.long.branch/dwell/out.json"),
    ("jc_long_trans", "trans", "JC.long.branch/trans/in.json", "JC.long.branch/trans/out.json"),
    # JC.long.branch/README.md:29-57,82-85
    ("jc29_same_ll", "ll", "JC.long.branch/jc29.same.json", {"columns": ["site", "value"], "data": [[0, -2.7725887222397811]]}),
    ("jc29_diff_ll", "ll", "JC.long.branch/jc29.diff.json", {"columns": ["site", "value"], "data": [[0, -2.7725887222397811]]}),
    ("jc30_same_ll", "ll", "JC.long.branch/jc30.same.json", {"columns": ["site", "value"], "data": [[0, -2.7725887222397811]]}),
    ("jc30_diff_ll", "ll", "JC.long.branch/jc30.diff.json", {"columns": ["site", "value"], "data": [[0, -2.7725887222397811]]}),
    ("jc29_same_deriv", "deriv", "JC.long.branch/jc29.same.json", {"columns": ["site", "edge", "value"], "data": [[0, 0, -6.4467380574161446e-17]]}),
    ("jc29_diff_deriv", "deriv", "JC.long.branch/jc29.diff.json", {"columns": ["site", "edge", "value"], "data": [[0, 0, 2.1489126858053815e-17]]}),
    ("jc30_same_deriv", "deriv", "JC.long.branch/jc30.same.json", {"columns": ["site", "edge", "value"], "data": [[0, 0, -1.6993417021166355e-17]]}),
    ("jc30_diff_deriv", "deriv", "JC.long.branch/jc30.diff.json", {"columns": ["site", "edge", "value"], "data": [[0, 0, 5.6644723403887852e-18]]}),
    ("jc600_same_deriv", "deriv", "JC.long.branch/jc600.same.json", {"columns": ["site", "edge", "value"], "data": [[0, 0, 0.0]]}),
    ("bpp_ll", "ll", "bpp.phyl/ll/in.json", "bpp.phyl/ll/out.json"),
    ("bpp_deriv", "deriv", "bpp.phyl/deriv/in.json", "bpp.phyl/deriv/out.json"),
    # GeLL.test.likelihood/README.md:29-31, GeLL.driver.DNA/README.md:30-32
    ("gell_test", "ll", "GeLL.test.likelihood/in.json", {"columns": ["value"], "data": [[-2616.073919844292]]}),
    ("gell_driver", "ll", "GeLL.driver.DNA/in.json", {"columns": ["value"], "data": [[-2616.0735881244163]]}),
    ("fels_em_full", "em_update", "Felsenstein.2004.fig.16.4/em-update/with.full.data/in.json", "Felsenstein.2004.fig.16.4/em-update/with.full.data/out.json"),
    ("fels_em_leaf", "em_update", "Felsenstein.2004.fig.16.4/em-update/with.leaf.data/in.json", "Felsenstein.2004.fig.16.4/em-update/with.leaf.data/out.json"),
    ("fels_em_none", "em_update", "Felsenstein.2004.fig.16.4/em-update/with.no.data/in.json", "Felsenstein.2004.fig.16.4/em-update/with.no.data/out.json"),
    ("fels_hess_full", "hess", "Felsenstein.2004.fig.16.4/hess/with.full.data/in.json", "Felsenstein.2004.fig.16.4/hess/with.full.data/out.json"),
    ("fels_hess_leaf", "hess", "Felsenstein.2004.fig.16.4/hess/with.leaf.data/in.json", "Felsenstein.2004.fig.16.4/hess/with.leaf.data/out.json"),
    ("fels_hess_none", "hess", "Felsenstein.2004.fig.16.4/hess/with.no.data/in.json", "Felsenstein.2004.fig.16.4/hess/with.no.data/out.json"),
]


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    ex = os.path.join(ref, "examples")
    manifest = []
    for name, prog, inp, out in CASES:
        with open(os.path.join(ex, inp)) as f:
            jin = json.load(f)
        if isinstance(out, str):
            with open(os.path.join(ex, out)) as f:
                jout = json.load(f)
            src_out = "examples/" + out
        else:
            jout = out
            src_out = "README value, examples/" + os.path.dirname(inp) + "/README.md"
        with open(os.path.join(HERE, name + ".in.json"), "w") as f:
            json.dump(jin, f, separators=(",", ":"))
            f.write("\n")
        with open(os.path.join(HERE, name + ".out.json"), "w") as f:
            json.dump(jout, f)
            f.write("\n")
        manifest.append({"name": name, "program": prog,
                         "source_in": "examples/" + inp, "source_out": src_out})
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
        f.write("\n")
    print("wrote %d golden cases to %s" % (len(manifest), HERE))


if __name__ == "__main__":
    main()
