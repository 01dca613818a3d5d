/* placeholder header until the backtracking engine lands (oracle part 2) */
