"""B200-native pairwise-individual similarity path of tidypopgen (KING / IBS / allele sharing / GRM)."""
