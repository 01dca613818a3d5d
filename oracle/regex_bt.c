/* oracle part 2 (regex engine) is added in a later commit */
