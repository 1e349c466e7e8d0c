"""TEST INFRASTRUCTURE: CPU oracle (C restatement + native build of the reference). Not imported by aim_b200."""
