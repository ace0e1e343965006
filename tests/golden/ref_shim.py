"""Import shim for the UNMODIFIED reference (build container: /root/reference; elsewhere the staged copy oracle/_ref).
The implementation lives in oracle/ref_loader.py; this module keeps the names make_golden.py and the reference-gated
tests use."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_loader  # noqa: E402

REFERENCE_ROOT = ref_loader.reference_root() or "/root/reference"
CfgNode = ref_loader.CfgNode
hrnet_config = ref_loader.hrnet_config
load_reference = ref_loader.load_reference
_stub = ref_loader._stub
