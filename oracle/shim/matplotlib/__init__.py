"""TEST INFRASTRUCTURE ONLY (oracle shim): empty matplotlib stub."""
