"""Drop-in counterparts of the reference's ``models/`` package (same module and class names)."""
