"""Seeded synthetic inputs (the generators live in scda_b200/synthetic.py so that bench.py
does not depend on the test tree)."""
from scda_b200.synthetic import *  # noqa: F401,F403
from scda_b200.synthetic import load_cfg  # noqa: F401
