"""centernet-lightning_b200: sm_100a CenterNet inference path behind the reference's Python API."""
__version__ = "0.1.0"
