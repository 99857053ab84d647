"""Import stub for nvalchemi-toolkit-ops 0.4.0 (absent offline). TEST INFRASTRUCTURE ONLY."""
__version__ = "0.4.0-stub"
