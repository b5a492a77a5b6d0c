"""TEST INFRASTRUCTURE ONLY -- see oracle/metro_oracle.py."""
