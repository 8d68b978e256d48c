"""Test infrastructure: CPU oracle of Crab's hot path (see crab_oracle.py). Not part of the product."""
