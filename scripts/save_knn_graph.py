"""`python save_knn_graph.py ...` as invoked by the reference's MID-FC/run_save_knn.py:50-66 (the script is missing
from the reference repository); see csn_b200/knn_driver.py."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from csn_b200.knn_driver import main  # noqa: E402

if __name__ == "__main__":
    raise SystemExit(main())
