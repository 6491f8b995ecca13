"""CPU oracle for the JARVIS-HybridNet 3D hot path — TEST INFRASTRUCTURE, not product code."""
