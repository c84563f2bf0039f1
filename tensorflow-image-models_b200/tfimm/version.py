# API level of the tfimm release this engine mirrors (reference: tfimm/version.py), plus our own.
__version__ = "0.2.14+b200.0.1"
