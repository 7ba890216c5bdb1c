"""Drop-in for the reference's ``src/network`` package (same module and class names)."""
