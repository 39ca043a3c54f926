"""Import-path mirrors of the reference's demo utilities that sit on the hot path (demos/<name>/utils/...)."""
