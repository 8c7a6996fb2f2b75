"""CPU oracle of the ClipCap hot path. TEST INFRASTRUCTURE ONLY — see oracle/restate.py for the import rules."""
