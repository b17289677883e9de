"""Extract the parameter-rule literals of the reference's known-answer tests into a JSON fixture.

    python tests/golden/make_kat_rules.py     (build container: reads /root/reference/tests)

``test_codon_rate_het`` / ``test_time_rate_het`` / ``test_loci``
(/root/reference/tests/test_evolve/test_likelihood_function.py:1522-2188) pin lnL values for
likelihood functions configured by long ``rules = [...]`` literals; the literals are data, lifted
here with ``ast`` so that tests/dropin_cases.py can apply them on the GPU box (where
/root/reference does not exist)."""

import ast
import json
import os

SRC = "/root/reference/tests/test_evolve/test_likelihood_function.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_rules.json")
WANTED = ("test_codon_rate_het", "test_time_rate_het", "test_loci")


def main():
    tree = ast.parse(open(SRC).read())
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in WANTED:
            for stmt in node.body:
                if (isinstance(stmt, ast.Assign) and len(stmt.targets) == 1
                        and getattr(stmt.targets[0], "id", None) == "rules" and node.name not in out):  # fmt: skip
                    out[node.name] = ast.literal_eval(stmt.value)
    assert set(out) == set(WANTED), sorted(out)
    with open(OUT, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print({k: len(v) for k, v in out.items()}, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
