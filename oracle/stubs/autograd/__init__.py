"""Import-only placeholder; autograd.numpy is plain numpy here."""
