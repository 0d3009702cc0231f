"""Parameter-holding mirrors of the reference's src/modules/* classes (same names, same state_dict keys)."""
