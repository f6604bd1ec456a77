"""CPU oracle of the mp2p_icp Matcher+Solver hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``mp2p_icp_b200/`` may import this package; see oracle/oracle.cpp.
"""
