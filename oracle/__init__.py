"""CPU oracle for the salva3d fluid-step path (TEST INFRASTRUCTURE ONLY; see oracle/oracle.cpp header).

Nothing under salva_b200/ may import this package.
"""
