"""Architecture-name coverage against a reference model registry.

    python scripts/registry_coverage.py /path/to/reference/vllm/model_executor/models/registry.py

Counts the architecture strings of the reference registry that this package resolves: natively registered classes,
pooling adapters derived from them (``XModel`` / ``XForSequenceClassification`` / reward heads), and speculative-decoding
head architectures served by the proposers in ``spec_decode``.  Prints the totals and the names still missing.
"""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main() -> None:
    path = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/vllm/model_executor/models/registry.py"
    ref = set(re.findall(r'^\s+"(\w+)":\s*\(', open(path).read(), re.M))
    from vllm_b200.models import registry as r
    from vllm_b200.spec_decode import DRAFT_ARCHS
    native = {a for a in ref if a in r._MODELS}
    adapted = {a for a in ref - native if r._adapter_for(a)}
    drafts = {a for a in ref - native - adapted if a in DRAFT_ARCHS}
    missing = sorted(ref - native - adapted - drafts)
    print(f"reference architectures: {len(ref)}")
    print(f"  native classes       : {len(native)}")
    print(f"  pooling adapters     : {len(adapted)}")
    print(f"  draft / MTP heads    : {len(drafts)}")
    print(f"  missing              : {len(missing)}")
    print("registered here (incl. names the reference does not list):", len(r._MODELS))
    if "-v" in sys.argv:
        print("\n".join(missing))


if __name__ == "__main__":
    main()
