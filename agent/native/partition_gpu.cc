// placeholder compiled by the Makefile until the real tool lands in the next commit
int main() { return 0; }
