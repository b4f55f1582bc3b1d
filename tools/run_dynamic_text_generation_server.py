#!/usr/bin/env python
"""Text-generation server on the continuous-batching engine (reference ``tools/run_dynamic_text_generation_server.py``): same REST surface as
``run_text_generation_server.py`` with ``--engine dynamic`` (paged KV cache, requests join and leave between decode steps)."""
import os
import runpy
import sys

if __name__ == "__main__":
    if "--engine" not in sys.argv:
        sys.argv += ["--engine", "dynamic"]
    runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "run_text_generation_server.py"), run_name="__main__")
