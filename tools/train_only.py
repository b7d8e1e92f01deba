"""The hypergraph2vec training extra of bench.py alone (for ncu captures of k_hg2v_epoch)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import hg2v_train_extra  # noqa: E402
from hypergraphembedding_b200 import _native  # noqa: E402

print(json.dumps(hg2v_train_extra(_native.default_context(0), epochs=int(sys.argv[1]) if len(sys.argv) > 1 else 1)))
