"""CPU oracle for the llm.f90 decode path -- TEST INFRASTRUCTURE ONLY (see llama2_oracle.c).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm.
The product package (llm.f90_b200) never imports this.
"""
