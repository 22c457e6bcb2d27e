"""Test infrastructure only: ctypes access to the plain-C restatement (oracle/rfq_oracle.c) and to the
unmodified reference binary (oracle/_ref/repaq).  Nothing under repaq_b200/ imports this package."""
