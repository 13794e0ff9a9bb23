"""ORACLE package — test infrastructure only (see dgcnn_oracle.py). Never imported by the product."""
