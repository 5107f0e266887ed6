"""`python -m pistachio_b200 structure.yaml out/ -p|-s`: the reference's command line (transfer_matrix.py:755-805)."""
from .transfer_matrix import main, parse_arguments

if __name__ == "__main__":
    main(parse_arguments())
