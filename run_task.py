"""Drop-in entry point: `python3 run_task.py <cfg.yml>` (reference: run_task.py:155-160)."""
import argparse

import vlb200  # noqa: F401
from vlb200.run_task import main

if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("init_file", help="Configuration .yml file for the run.")
    parser.add_argument("--device", default="cuda:0")
    args = parser.parse_args()
    main(args.init_file, args.device)
