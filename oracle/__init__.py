"""TEST INFRASTRUCTURE ONLY: CPU oracle for the hot path. See oracle/oracle.cpp header."""
